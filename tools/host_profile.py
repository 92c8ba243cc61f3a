"""Host-side cost of queue_frame on a small frame (config 1: 640x360 G3, 256 spp), where the
GPU work is shorter than the Python in front of it: cProfile of 200 pipelined frames."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuburn_b200 import _native as N, samples, profile, render

N.init(0)
gnm = samples.g3()
gprof = profile.wrap(dict(width=640, height=360, spp=256, frame_width=0, start=1, end=2), gnm)
tc = profile.enumerate_times(gprof)[0][1][0]
rmgr = render.RenderManager(seed=1)
rdr = render.Renderer(gnm, gprof)
for _ in range(5):
    evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc)
    evt.synchronize()


def run(n):
    pending = []
    for k in range(n):
        pending.append(rmgr.queue_frame(rdr, gnm, gprof, tc))
        if len(pending) > 2:
            pending.pop(0)[0].synchronize()
    for evt, _ in pending:
        evt.synchronize()


t = time.perf_counter(); run(200); dt = time.perf_counter() - t
print('pipelined: %.3f ms per frame (%.0f frames/s)' % (dt / 200 * 1e3, 200 / dt))
evt, buf = rmgr.queue_frame(rdr, gnm, gprof, tc); evt.synchronize()
print('GPU time of one frame: %.3f ms' % evt.time())
pr = cProfile.Profile(); pr.enable(); run(200); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:6000])
