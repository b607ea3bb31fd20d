"""GPU-box debug aid: column-edge pass alone and row-edge pass alone (RB200_LF_ONLY_DIR) against a numpy port."""
import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] in ("0", "1"):
    only = int(sys.argv[1])
    os.environ["RB200_LF_ONLY_DIR"] = str(only)
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import streamdump
    from rav1d_b200 import lib
    from rav1d_b200.synth import framegen
    lib.check(lib.init(0))
    key = "8-bit/data/00000658.ivf#0"
    s = dict(streamdump.load_golden())[key]
    pre = s.pre
    if len(sys.argv) > 2:      # rows pass alone on the numpy port's column-pass output
        luma = np.load(sys.argv[2])
        pre = [p.copy() for p in s.pre]
        pre[0][:luma.shape[0], :luma.shape[1]] = luma
    d = framegen.DeviceFrame(s); d.load_batch(); d.upload(0, pre); d.submit(2); d.wait()
    got = streamdump.visible(s, d.readback()); d.close()
    np.save(f"/tmp/dbg_lf_only{only}.npy", got[0])
else:
    for only in (0, 1):
        subprocess.check_call([sys.executable, __file__, str(only)])
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import streamdump
    s = dict(streamdump.load_golden())["8-bit/data/00000658.ivf#0"]
    sys.argv = ["x"]
    ns = {}
    src = open(os.path.join(ROOT, "tools", "lfsim.py")).read()
    for only in (0, 1):
        ns = {"ONLY": only}
        exec(compile(src.replace("for d in (0,1):", "for d in (ONLY,):"), "lfsim", "exec"), ns)
        exp = ns["vis"]; got = np.load(f"/tmp/dbg_lf_only{only}.npy")
        bad = np.argwhere(exp != got)
        print("pass", only, "alone:", len(bad), "px differ", [(int(x), int(y), int(exp[y, x]), int(got[y, x]), int(s.pre[0][y, x])) for y, x in bad[:16]])
        if only == 0: np.save("/tmp/dbg_cols_full.npy", ns["pic"].astype(s.pre[0].dtype))
    subprocess.check_call([sys.executable, __file__, "1", "/tmp/dbg_cols_full.npy"])
    ns = {}
    exec(compile(src, "lfsim", "exec"), ns)
    exp = ns["vis"]; got = np.load("/tmp/dbg_lf_only1.npy")
    bad = np.argwhere(exp != got)
    print("rows pass on the port's column-pass output:", len(bad), "px differ", [(int(x), int(y), int(exp[y, x]), int(got[y, x])) for y, x in bad[:16]])
