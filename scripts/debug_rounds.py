"""Diagnostics (build with -DFS2_DEBUG_ROUNDS, FS2_LIB=.../libfs2_dbg.so): per step of the bench stream, the share of
particles whose observations needed more than one speculative round, or the sequential fall-back."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fast_slam_b200 import _lib
from fast_slam_b200.filter import _hash_uniform

P = 1 << 20
flt, world = bench.make_synthetic_filter(P, bench.L, bench.LCAP)
was = False
for s in range(int(os.environ.get("FS2_DBG_STEPS", "26"))):
    rot, tr, obs = bench.synthetic_step_inputs(bench.SEED, s, world, bench.M)
    flt.status.zero_()
    flt.draw_noise(0.001 if rot != 0 else 0.0055, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); flt.motion_update(rot, tr, obs); e1.record()
    flt.finish_step(_hash_uniform(bench.SEED, s) / P)          # the bench's step: resample decided on the device, copies deferred
    stats = flt.stats.cpu()
    st = flt.status
    torch.cuda.synchronize()
    sh = lambda bit: float((st & bit).ne(0).double().mean())
    print("step %2d%s %.2f ms  seq %.5f  >1 round %.4f  >2 %.4f  >4 %.4f  mean count %.1f max %d | first dependency: same landmark %.3f, "
          "captured by new state %.3f; dependent obs unmatched %.3f / matched %.3f | resampled %d deferred %d" % (
              s, "+" if was else " ", e0.elapsed_time(e1), sh(16), sh(32), sh(64), sh(128), float(flt.count.double().mean()), int(flt.count.max()),
              sh(256), sh(512), sh(1024), sh(2048), int(stats[_lib.STAT_RESAMPLED]), int(stats[_lib.STAT_DEFERRED])), flush=True)
    was = bool(stats[_lib.STAT_DEFERRED] > 0)
