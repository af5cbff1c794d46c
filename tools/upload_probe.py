"""Host frame -> device upload of a crop's rectangle: the previous path (rows into a pinned buffer with torch, one contiguous H2D) against
vt_upload_frame_rect (the rectangle packed by a few threads, one strided H2D).  Wall clock per call incl. stream sync.
    python tools/upload_probe.py"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from vittracker_b200 import _lib
    lib = _lib.load()
    H, W = 720, 1280
    img = np.random.default_rng(0).integers(0, 255, size=(H, W, 3), dtype=np.uint8)
    dev = torch.zeros(H * W * 3, dtype=torch.uint8, device="cuda")
    pin = torch.empty(H * W * 3, dtype=torch.uint8).pin_memory()
    stage = torch.empty(H * W * 3, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    def t(fn, n=400):
        for _ in range(20): fn()
        xs = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); st.synchronize(); xs.append(time.perf_counter() - t0)
        return 1e6 * float(np.median(xs))
    for (ya, yb, xa, xb) in ((20, 700, 200, 944), (100, 458, 300, 658), (0, 720, 0, 1280), (300, 400, 500, 600)):
        a, b = ya * W * 3, yb * W * 3
        def old():
            pin[a:b].copy_(torch.from_numpy(img.reshape(-1)[a:b]))
            dev[a:b].copy_(pin[a:b], non_blocking=True)
        def new():
            lib.vt_upload_frame_rect(img.ctypes.data, H, W, ya, yb, xa, xb, pin.data_ptr(), stage.data_ptr(), dev.data_ptr(), st.cuda_stream)
        def new2d():
            lib.vt_upload_frame_rect(img.ctypes.data, H, W, ya, yb, xa, xb, pin.data_ptr(), None, dev.data_ptr(), st.cuda_stream)
        print(f"  rect rows {yb - ya} x cols {xb - xa} ({(yb - ya) * (xb - xa) * 3 / 1e6:.2f} MB of {(b - a) / 1e6:.2f} MB rows): rows+torch {t(old):7.1f} us   rect {t(new):7.1f} us   rect, strided H2D {t(new2d):7.1f} us", flush=True)
    sys.exit(0)
subprocess.run([sys.executable, __file__, "child"])
