"""Where the host-buffer iteration's time goes: bare pinned copies against the whole m3dreg_icp_iteration_host call.
Run twice (M3DREG_HOST_OVERLAP=0 / 1) to see what running the grid build under the second upload buys."""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("mandala-mapping_b200")
first, second, pose_init, pose2, pose_true = pkg.synth.scan_pair("sick", seed=42)
n1, n2 = len(first), len(second)
prm = pkg.default_params(1.0, dof=6)
ctx = pkg.Context(0)
h_first = torch.from_numpy(np.frombuffer(first.tobytes(), dtype=np.uint8).copy()).pin_memory()
h_second = torch.from_numpy(np.frombuffer(second.tobytes(), dtype=np.uint8).copy()).pin_memory()
h_nn = torch.empty(n2, dtype=torch.int32).pin_memory()
d = torch.empty(h_first.numel() + h_second.numel(), dtype=torch.uint8, device="cuda")
dn = torch.empty(n2, dtype=torch.int32, device="cuda")
def t(fn, k=10):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
def copies():
    d[:h_first.numel()].copy_(h_first, non_blocking=True); d[h_first.numel():].copy_(h_second, non_blocking=True)
pose = np.ascontiguousarray(pose_init, dtype=np.float32).reshape(16).copy()
print("overlap env:", os.environ.get("M3DREG_HOST_OVERLAP", "(default 1)"))
print("bare H2D of both clouds: %.3f ms" % t(copies))
print("bare D2H of nn: %.3f ms" % t(lambda: h_nn.copy_(dn, non_blocking=True)))
print("m3dreg_icp_iteration_host (nn out): %.3f ms" % t(lambda: bench._e2e_call(ctx, pkg, h_first, n1, h_second, n2, pose, prm, h_nn)))
st = bench._e2e_call(ctx, pkg, h_first, n1, h_second, n2, pose, prm, h_nn)
print("device_ms of the iteration inside the call: %.3f" % st.device_ms)
if os.environ.get("M3DREG_HOST_TRACE"):
    print("-- one traced call (stderr)")
    bench._e2e_call(ctx, pkg, h_first, n1, h_second, n2, pose, prm, h_nn)
