import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
from fers_core_b200 import engine, synth
big, _ = synth.portal_batch(1024, seed=synth.BATCH_SEED)
dev = torch.device("cuda", 0)
for _ in range(2):
    engine.analyze_frame_batch(big, synth.PORTAL_NODES, dev)
pr = cProfile.Profile()
pr.enable()
engine.analyze_frame_batch(big, synth.PORTAL_NODES, dev)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
