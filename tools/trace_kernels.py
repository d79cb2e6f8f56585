"""In-kernel timelines of the convgemm launches of one transformer layer at config-2 size (debug build).

    bash speech_resynth_b200/csrc/build.sh trace     # -> build/libsrb_trace.so (-DSRB_TRACE)
    SRB_DEBUG_LIB=build/libsrb_trace.so python tools/trace_kernels.py

Prints, for the first CTAs of each launch, %globaltimer stamps relative to the CTA's entry (microseconds):
producer: entry, prologue done, dependencies satisfied, finished; MMA warp per tile: accumulator free, first box
landed, MMAs issued; epilogue warp 2 per tile: ready, accumulator complete, epilogue done.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import speech_resynth_b200 as srb  # noqa: E402
from speech_resynth_b200 import _native as nat  # noqa: E402
from speech_resynth_b200 import engine as eng  # noqa: E402
from speech_resynth_b200 import synthetic  # noqa: E402

P = nat.ptr


def dump(trace, title):
    t = trace.cpu().view(8, 7, 8, 4).tolist()
    print("==", title)
    for cta in (0, 1, 147 % 8):
        t0 = t[cta][0][0][0]
        if t0 == 0:
            continue
        rel = lambda v: f"{(v - t0) / 1e3:7.2f}" if v else "      -"
        print(f" cta {cta}: producer entry 0, prologue {rel(t[cta][0][0][1])}, deps {rel(t[cta][0][0][2])}, done {rel(t[cta][0][0][3])}")
        for it in range(8):
            if t[cta][1][it][0] == 0 and t[cta][2][it][0] == 0:
                break
            print(f"   tile {it}: mma free {rel(t[cta][1][it][0])} box {rel(t[cta][1][it][1])} issued {rel(t[cta][1][it][2])} | "
                  f"epi ready {rel(t[cta][2][it][0])} acc {rel(t[cta][2][it][1])} done {rel(t[cta][2][it][2])}")
            if t[cta][3][it][0]:
                # RESNORM epilogue of warp 2, per 32-column chunk: residual arrived / fp32 block stored / bf16 block stored
                print("       resid " + " ".join(rel(v) for v in t[cta][3][it]) + " | fp32 " + " ".join(rel(v) for v in t[cta][4][it]) +
                      " | bf16 " + " ".join(rel(v) for v in t[cta][5][it]) + f" | acc0 {rel(t[cta][6][it][0])} norm {rel(t[cta][6][it][1])}")


if __name__ == "__main__":
    lib = nat.load()
    lib.srb_debug_set_trace.argtypes = [ctypes.c_void_p]
    lib.srb_debug_set_trace.restype = None
    b, n = 64, 500
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.cuda()
    e = eng.ResynthEngine(decoder.model.sampler(), decoder.vocoder.generator(), use_graphs=False)
    sampler = e.sampler
    n8 = eng.padded_frames(n)
    ws = sampler.workspace(b, n8, mel_rows=n)
    ws["xn"].zero_()
    g = sampler.cond_table(eng.ode_times(0.0625))
    sampler.rotary(n8)
    sampler.fork.enabled = False
    sampler.stage(ws, synthetic.make_units(b, n, seed=7).cuda(), torch.randn(b, n, 80, device="cuda"), 1.0)
    sampler.prepare(ws)
    sampler.step(ws, g[0], 0.0625, last=False)     # warm: fills every buffer with real data
    torch.cuda.synchronize()
    trace = torch.zeros(8 * 7 * 8 * 4, dtype=torch.int64, device="cuda")
    w, L = sampler.w, ws["lengths"]
    cs, sn = sampler.rotary(n8)
    m = b * n8
    calls = {
        "qk_rope": lambda: nat.call("srb_cfm_qk_rope", P(ws["xn"]), P(w.w_qkv[0]), P(cs), P(sn), P(ws["qk"]), P(ws["qkmax"][0]), None, b, n8),
        "v_transposed": lambda: nat.call("srb_cfm_v_transposed", P(ws["xn"]), P(w.w_qkv[0][512:]), P(ws["vt"]), ws["vt"].shape[1]),
        "attn_out_norm": lambda: nat.call("srb_cfm_attn_out_norm", P(ws["o"]), P(w.w_out[0]), P(g[0][1]), P(L), P(ws["x"]), P(ws["xn"]), b, n8),
        "ffn_glu": lambda: nat.call("srb_cfm_ffn_glu", P(ws["xn"]), P(w.w_ff1[0]), P(w.b_ff1[0]), P(L), P(ws["h"]), b, n8, 1),
        "ffn_out_norm": lambda: nat.call("srb_cfm_ffn_out_norm", P(ws["h"]), P(w.w_ff2[0]), P(w.b_ff2[0]), P(g[0][2]), 1, P(L), P(ws["x"]), P(ws["xn"]), b, n8),
        "embed": lambda: nat.call("srb_cfm_embed", P(ws["xt_b"]), P(w.w_embed), P(ws["cond"]), P(ws["x0"]), b, n8),
        "pred_euler": lambda: nat.call("srb_cfm_pred_euler", P(ws["xn"]), P(w.w_pred), 0.0625, P(ws["xt"]), P(ws["xt_b"]), None, None, n, 2.26, -5.88, -11.5, P(L), b, n8),
    }
    # attention: raw ordered stamps per role (see ATTN_STAMP in csrc/srb_attention_tc.cu)
    att = torch.zeros(4 * 3 * 256, dtype=torch.int64, device="cuda")
    lib.srb_debug_set_trace(att.data_ptr())
    nat.call("srb_cfm_attention_tc", P(ws["qk"]), 512, P(ws["vt"]), ws["vt"].shape[1], P(L), P(ws["qkmax"][0]), P(ws["o"]), b, n8)
    torch.cuda.synchronize()
    lib.srb_debug_set_trace(None)
    a = att.cpu().view(4, 3, 256)
    t0 = int(a[0, 0, 0])
    for role, nm in enumerate(("producer", "mma", "softmax warp 2")):
        vals = [int(v) for v in a[0, role] if int(v) != 0]
        print(f"== attention cta 0 {nm}: " + " ".join(f"{(v - t0) / 1e3:.2f}" for v in vals))
    for name, fn in calls.items():
        trace.zero_()
        torch.cuda.synchronize()
        lib.srb_debug_set_trace(trace.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        lib.srb_debug_set_trace(None)
        dump(trace, f"{name}  ({e0.elapsed_time(e1) * 1e3:.1f} us between events)")
