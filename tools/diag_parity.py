"""GPU diagnostic: stage-by-stage error of the CUDA path against the CPU oracle (test tooling)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import multimodal_vqvae_compression_audio_tactile_b200 as pkg
from oracle import cases, proposed

def rel(a, b):
    a = a.double(); b = b.double()
    return ((a - b).abs().max().item(), (a - b).norm().item() / max(b.norm().item(), 1e-30))

def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "f32"
    name = sys.argv[2] if len(sys.argv) > 2 else "c3_b10k128"
    case = cases.CODEC_CASES[name]
    dev = torch.device("cuda", 0)
    print("device", torch.cuda.get_device_name(0), "precision", prec, "case", name, flush=True)
    t0 = time.time()
    ref = cases.build_reference_style_model(proposed.ProposedEval, case)
    a, t = cases.codec_inputs(case)
    tr = {}
    y_ref = ref.forward_eval(a, t, case.get("books_use"), trace=tr)
    print(f"oracle forward {time.time()-t0:.1f}s", flush=True)
    net = pkg.build_proposed(case["books"], case["K"])
    net.load_state_dict(ref.state_dict())
    for m in (net, net.A_ENC, net.T_ENC, net.T_DEC, net.A_QUANT, net.predict, net.vq):
        m.precision = prec
    ad, td = a.to(dev), t.to(dev)
    # stage by stage, teacher-forced with oracle tensors
    za = net.A_ENC(ad).cpu();           print("A_ENC   max/rel", rel(za, tr["za"]), flush=True)
    zt = net.T_ENC(td).cpu();           print("T_ENC   max/rel", rel(zt, tr["zt"]), flush=True)
    qa, codes, *_ = net.A_QUANT(tr["za"].to(dev))
    print("A_QUANT max/rel", rel(qa.cpu(), tr["qa"]), "code mismatches", int((codes.cpu() != tr["a_codes"]).sum()), "/", codes.numel(), flush=True)
    y_tf = net.T_DEC(tr["z_run"].to(dev)).cpu(); print("T_DEC   max/rel", rel(y_tf, y_ref), flush=True)
    zp, zk = cases.predictor_inputs()
    pr_ref = ref.predict(zp, zk)
    print("predict max/rel", rel(net.predict(zp.to(dev), zk.to(dev)).cpu(), pr_ref), flush=True)
    rd = tr["rD"]
    q_ref = ref.vq(rd, case.get("books_use")); i_ref = ref.vq.last_indices
    q_gpu, i_gpu = net.vq(rd.to(dev), case.get("books_use"), return_indices=True)
    print("vq      max/rel", rel(q_gpu.cpu(), q_ref), "idx mismatches", int((i_gpu.cpu() != i_ref).sum()), "/", i_ref.numel(), flush=True)
    # end to end
    torch.cuda.synchronize(); t1 = time.time()
    y = net.forward_eval(ad, td, case.get("books_use")); torch.cuda.synchronize()
    print(f"fused forward (first, incl. pack) {time.time()-t1:.2f}s")
    idx = net.last_indices.cpu().long()
    mism = (idx != tr["idx"])
    print("E2E y   max/rel", rel(y.cpu(), y_ref), "idx mismatches", int(mism.sum()), "/", idx.numel(),
          "margins at mismatches", tr["margin"][mism][:8].tolist(), flush=True)
    print("E2E audio codes mismatches", int((net.last_audio_codes.cpu().long() != tr["a_codes"]).sum()))
    z = net.encode_latents(ad, td, case.get("books_use")).cpu()
    print("z_run   max/rel", rel(z, tr["z_run"]))
    # timing
    for B in (1, 8, 32):
        ab = torch.rand(B, 1, 24000, device=dev) * 2 - 1; tb = torch.rand(B, 1, 24000, device=dev) * 2 - 1
        net.forward_eval(ab, tb); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); n = 3
        for _ in range(n): net.forward_eval(ab, tb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"B={B}: {ms:.2f} ms/forward -> {B / ms * 1e3:.1f} signal-s/s", flush=True)

if __name__ == "__main__":
    main()
