mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_misfit_metrics.py -x -q --timeout 60 2>&1 | tail -15
timeout 100 python - <<'PY' 2>&1 | tee gpurun_out/wasserstein_bench.log
import time, numpy as np, torch, ertdiff_b200 as eb
from scipy.stats import wasserstein_distance
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
obs = rng.normal(size=(4693, 14)).astype(np.float32)
sims = obs[None] + rng.normal(scale=0.2, size=(50, 4693, 14)).astype(np.float32)
ts, to = torch.from_numpy(sims).to(dev), torch.from_numpy(obs).to(dev)
for N in (1, 50):
    for _ in range(2): eb.wasserstein_distance(ts[:N], to)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d = eb.wasserstein_distance(ts[:N], to)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter(); ref = [wasserstein_distance(sims[i].flatten(), obs.flatten()) for i in range(min(N, 5))]
    cpu = (time.perf_counter() - t0) / min(N, 5)
    print(f"N {N}: {e0.elapsed_time(e1) / 5 * 1e3:.0f} us per call on the device; scipy {cpu * 1e3:.1f} ms per pair; max rel err {max(abs(d[i].item() - ref[i]) / ref[i] for i in range(len(ref))):.2e}")
PY
