import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ertdiff_b200 as eb
dev = torch.device("cuda", 0)
N = int(os.environ.get("MEMBERS", "256"))
torch.manual_seed(0)
model = eb.ConditionalDiffusionModel(29, 128).to(dev).eval()
cond = torch.rand(1, 14, 4693, device=dev).expand(N, 14, 4693)
sched = [t.to(dev) for t in eb.get_diffusion_schedule(1000)]
x = eb.run_chain(model, cond, 1000, *sched, dev, seed=1)
for _ in range(3):
    m = eb.ensemble_kde_mode(x, 5000)
torch.cuda.synchronize()
print("ok", m[:3])
