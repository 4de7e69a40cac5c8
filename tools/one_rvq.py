"""One residual-VQ forward (the codec path's shapes) for ncu: python tools/one_rvq.py [B] [T] [books] [K] [D]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_vqvae_compression_audio_tactile_b200 as pkg

B, T, books, K, D = (int(a) for a in (sys.argv[1:6] + ["64", "75", "8", "512", "96"][len(sys.argv) - 1:]))
torch.manual_seed(0)
vq = pkg.ResidualVQEMA(dim=D, n_books=books, n_embed=K).cuda()
z = (torch.randn(B, D, T) / D ** 0.5).cuda()
for prec in ("tc", "f32"):
    vq.precision = prec
    for _ in range(3):
        q, idx = vq(z, return_indices=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        vq(z)
    e1.record()
    torch.cuda.synchronize()
    print(prec, "ms per forward (incl. transposes)", e0.elapsed_time(e1) / 10, int(idx.sum()))
