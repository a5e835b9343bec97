"""MSDeformAttn kernel variants side by side (C2: 3 levels N=2; C5(i): 4 levels N=1): the L1-resident tiled kernels of
round 1, the TMA-staged fp32 kernel and the opt-in fp16-staged kernel (gather alone and gather + repack), forward and
backward.  CUDA-event timing, one launch per event pair, L2 flushed ahead of every launch; algorithmic bytes as in
DESIGN.md 3 (fp32 value + loc + attn in, out).  Nothing here is a bench.py value."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from partdistillation_b200 import _lib, functional as fn  # noqa: E402
from tools.microbench import msda_case, timeit  # noqa: E402


def main():
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    lib = _lib.load()
    res = {}
    for name, N, shapes, spread in (("C2_3lvl_N2", 2, [(32, 32), (64, 64), (128, 128)], 4.0),
                                    ("C2_3lvl_N2_spread2", 2, [(32, 32), (64, 64), (128, 128)], 2.0),
                                    ("C5i_4lvl_N1", 1, [(256, 256), (128, 128), (64, 64), (32, 32)], 4.0)):
        value, loc, attn, shapes, fb, bb = msda_case(N, shapes, spread)
        S = value.shape[1]
        r = {"fwd_algorithmic_bytes": fb, "bwd_algorithmic_bytes": bb}
        with torch.no_grad():
            lib.pdb_debug_set_msda_path(1)
            t = timeit(lambda: fn.ms_deform_attn(value, shapes, None, loc, attn), flush=flush)
            ref = fn.ms_deform_attn(value, shapes, None, loc, attn)
            r["fwd_l1_tiled_us"] = t * 1e6
            lib.pdb_debug_set_msda_path(4)
            t = timeit(lambda: fn.ms_deform_attn(value, shapes, None, loc, attn), flush=flush)
            out = fn.ms_deform_attn(value, shapes, None, loc, attn)
            lib.pdb_debug_set_msda_path(0)
            r["fwd_tma_f32_us"] = t * 1e6
            r["fwd_tma_f32_frac"] = fb / t / 1e9 / peak
            r["fwd_tma_vs_l1_rel"] = float((out - ref).abs().max() / ref.abs().max())
            fn.msda_value_half = True
            t = timeit(lambda: fn.ms_deform_attn(value, shapes, None, loc, attn), flush=flush)
            outh = fn.ms_deform_attn(value, shapes, None, loc, attn)
            r["fwd_half_incl_repack_us"] = t * 1e6
            fn.msda_value_half = False
            vh = torch.empty((N, 8, S, 32), dtype=torch.float16, device="cuda")
            hs = _lib.host_i64([v for hw in shapes for v in hw])
            starts, s = [], 0
            for h, w in shapes:
                starts.append(s)
                s += h * w
            st = _lib.host_i64(starts)
            stream = torch.cuda.current_stream().cuda_stream
            t = timeit(lambda: lib.pdb_msda_pack_value_h(value.data_ptr(), vh.data_ptr(), N, S, 8, 32, stream), flush=flush)
            r["pack_value_h_us"] = t * 1e6
            o2 = torch.empty_like(out)
            t = timeit(lambda: lib.pdb_msda_forward_h(vh.data_ptr(), hs, st, loc.data_ptr(), attn.data_ptr(), o2.data_ptr(), N, S, 8,
                                                       32, S, len(shapes), 4, stream), flush=flush)
            r["fwd_half_gather_us"] = t * 1e6
            r["fwd_half_gather_frac"] = fb / t / 1e9 / peak
            lib.pdb_debug_set_msda_path(2)
            t = timeit(lambda: lib.pdb_msda_forward_h(vh.data_ptr(), hs, st, loc.data_ptr(), attn.data_ptr(), o2.data_ptr(), N, S, 8,
                                                       32, S, len(shapes), 4, stream), flush=flush)
            lib.pdb_debug_set_msda_path(0)
            r["fwd_half_gather_2cta_us"] = t * 1e6
            r["fwd_half_rel_err_vs_f32"] = float((outh - out).abs().max() / out.abs().max())
        out = fn.ms_deform_attn(value, shapes, None, loc, attn)
        go = torch.randn_like(out)
        tb = timeit(lambda: torch.autograd.grad(out, (value, loc, attn), go, retain_graph=True), flush=flush)
        r["bwd_us"] = tb * 1e6
        r["bwd_frac"] = bb / tb / 1e9 / peak
        res[name] = r
    print(json.dumps(res, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_msda.json", "w"), indent=1)


if __name__ == "__main__":
    main()
