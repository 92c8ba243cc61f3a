mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r02_gputests_final2.log 2>&1; tail -6 gpurun_out/r02_gputests_final2.log
python -c "
import __graft_entry__ as g
g.smoke()
" > gpurun_out/r02_smoke_final2.log 2>&1; tail -2 gpurun_out/r02_smoke_final2.log
