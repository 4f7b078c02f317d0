"""cProfile of JobScheduler.run on one GPU (where does a job's set-up time go: plan creation, cudaFree, pinned
allocations) -- the measurement that led to pooled plans and the pinned trace ring."""
import cProfile, pstats, sys, os
sys.path.insert(0, os.getcwd())
import bench
from style_transfer2_b200 import serving
from style_transfer2_b200.model import B200Model
import torch
content, style, _ = bench.load_images(512)
jobs = [serving.job_messages(512, content, style, bench.WEIGHTS, bench.PARAMS, seed=j) for j in range(6)]
model = B200Model(precision='fp16')
sched = serving.JobScheduler(model, max_resident=8)
sched.run(jobs[:1], 3, fetch_final=False)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
sched.run(jobs, 20)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
