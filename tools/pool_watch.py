"""Diagnostics: render frames of a bench workload in a thread; if one does not end within a few seconds, read the device
counters from this thread (chaos_debug_peek_counters) and print the orbit-pool state of every strand and shard."""
import ctypes as C, importlib, os, sys, threading, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
cu = importlib.import_module("chaos-ultra_b200")
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 300
wl = bench.WORKLOADS[w]
prov = cu.CudaFractalRendererProvider(device=0)
r = prov.getRenderer(wl["fractal"], False)
r.initializeRendering(wl["W"], wl["H"], None, cu.OUTPUT_HOST if os.environ.get("WATCH_HOST") else cu.OUTPUT_DEVICE)
r.setPartition(0, 1, 32)
m = bench.make_model(cu, wl)
state = {"frame": -1, "done": False}
def work():
    for f in range(frames):
        state["frame"] = f
        r.renderQuality(m)
    state["done"] = True
t = threading.Thread(target=work, daemon=True); t.start()
last, since = -1, time.time()
while not state["done"]:
    time.sleep(0.5)
    if state["frame"] != last: last, since = state["frame"], time.time()
    elif time.time() - since > 8:
        lib = r._lib
        buf = np.zeros(1 << 20, dtype=np.uint8)
        n = lib.chaos_debug_peek_counters(r._h, buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.size))
        print("frame %d does not end; counters block = %d bytes" % (last, n))
        for s in range(2):
            blk = buf[s * n:(s + 1) * n]
            head = blk[:24].view(np.uint32)
            print(" strand %d: next_tile %d next_b %d tail_b %d claimed_b %d n_exported %d next_export_item %d" % ((s,) + tuple(int(x) for x in head)))
            off = 32 + 24 + 4 * 2 * 40     # cursors, 3 x u64, two bucket arrays
            pool = blk[off:off + 2 * 128 * 16].view(np.uint32).reshape(2, 128, 4)
            for p in range(2):
                bad = [(k, int(pool[p, k, 0]), int(pool[p, k, 1]), int(pool[p, k, 2])) for k in range(128) if pool[p, k, 0] or pool[p, k, 1] != pool[p, k, 2]]
                print("  pool[%d] shards alive or not drained (shard, live, reserved, head): %s" % (p, bad[:40]))
        os._exit(3)
print("all %d frames ended" % frames)
