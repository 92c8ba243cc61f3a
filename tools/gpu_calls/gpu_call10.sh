mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q --durations=6) > gpurun_out/r02_gputests_8.log 2>&1; tail -12 gpurun_out/r02_gputests_8.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
