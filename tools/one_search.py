import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_vqvae_compression_audio_tactile_b200 as pkg
from oracle import cases
N, D, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x, emb = cases.search_inputs(N, D, K)
x, emb = x.cuda(), emb.cuda()
for _ in range(3):
    i = pkg.nearest_code(x, emb, precision="tc")
torch.cuda.synchronize()
print("ok", int(i.sum()))
