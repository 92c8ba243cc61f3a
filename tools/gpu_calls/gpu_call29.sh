for e in RED_POLICY=0 RED_POLICY=1; do echo "== $e 4K"; EXTRA=$e CASES=dynamic:1 W=3840 H=2160 SPP=1000 timeout 300 python tools/sched_bench.py G6F 2>&1 | cut -c40-200; done
for e in RED_POLICY=0 RED_POLICY=1; do echo "== $e 1080p"; EXTRA=$e CASES=dynamic:1 FW=0 timeout 300 python tools/sched_bench.py G6F G3 2>&1 | cut -c40-200; done
for e in RED_POLICY=0 RED_POLICY=1; do echo "== $e 8K"; EXTRA=$e CASES=dynamic:1 FW=0 W=7680 H=4320 SPP=250 timeout 300 python tools/sched_bench.py G24H 2>&1 | cut -c40-200; done
