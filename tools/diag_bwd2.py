"""Diagnostic: wn_tb_wgrad and wn_tb_block_bwd_data in isolation against float64 torch math on random inputs."""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-wavenet_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
import native
lib = native.lib()
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator().manual_seed(0)
CH = 256


def to_pair(x):                      # (B, L, C) fp32 cuda -> pair tensor, and the value the pair represents (float64)
    B, L, C = x.shape
    pair = torch.zeros(B, 2, C // 8, L, 8, device=dev, dtype=torch.bfloat16)
    native.check(lib.wn_pair_from_frames(x.data_ptr(), pair.data_ptr(), B, L, C, 0, st), "pair")
    hi = x.to(torch.bfloat16); lo = (x - hi.float()).to(torch.bfloat16)
    return pair, (hi.double() + lo.double())


def from_pair(pair, C, t0=0):
    B, _, _, L, _ = pair.shape
    out = torch.zeros(B, L, C, device=dev)
    native.check(lib.wn_frames_from_pair(pair.data_ptr(), out.data_ptr(), B, L, C, t0, st), "frames")
    return out


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


for (B, L, d, in_start, out_start, gs_out, ds_start) in [(2, 420, 4, 6, 10, 60, 270), (3, 700, 8, 10, 18, 120, 400), (3, 700, 8, 10, 18, 700, 400)]:
    have_dh = gs_out < L
    gz = max(out_start, min(gs_out, ds_start))
    id_start = max(out_start, gs_out)
    gs_in = max(in_start, min(id_start, gz - d))
    print(f"--- B={B} L={L} d={d} in={in_start} out={out_start} gs_out={gs_out} ds={ds_start} gz={gz} id={id_start} gs_in={gs_in}")
    rnd = lambda *s: (torch.randn(*s, generator=g)).to(dev)
    dh_out = rnd(B, L, CH) * 1e-3
    dh_out[:, :min(gs_out, L)] = 0
    dskip = rnd(B, L - ds_start, CH) * 1e-3
    f = torch.tanh(rnd(B, L, CH)); gg = torch.sigmoid(rnd(B, L, CH))
    h_in = rnd(B, L, CH); h_in[:, :in_start] = 0
    wf, wg = rnd(CH, CH, 2) * 0.05, rnd(CH, CH, 2) * 0.05
    wr, ws = rnd(CH, CH, 1) * 0.05, rnd(CH, CH, 1) * 0.05
    dh_pair, dh_q = to_pair(dh_out)
    ds_pair, ds_q = to_pair(dskip)
    hin_pair, hin_q = to_pair(h_in)
    fg = torch.cat([f.view(B, L, CH // 4, 4).permute(0, 2, 1, 3), gg.view(B, L, CH // 4, 4).permute(0, 2, 1, 3)], 1).contiguous()
    wb = torch.empty(lib.wn_tb_bwd_weight_bytes_per_layer(CH, 2), device=dev, dtype=torch.uint8)
    ptrs = torch.tensor([[wf.data_ptr(), wg.data_ptr(), 0, 0, wr.data_ptr(), ws.data_ptr(), 0, 0]], dtype=torch.int64, device=dev)
    native.check(lib.wn_tb_pack_all_bwd_weights(ptrs.data_ptr(), 1, CH, 2, wb.data_ptr(), st), "pack")
    dfg = torch.zeros(B, 2, 64, L, 8, device=dev, dtype=torch.bfloat16)
    zb = torch.zeros(B, 2, 32, L, 8, device=dev, dtype=torch.bfloat16)
    dh_in = torch.zeros(B, 2, 32, L, 8, device=dev, dtype=torch.bfloat16)
    a = native.TbBwdArgs()
    a.d_dh_out = dh_pair.data_ptr() if have_dh else None
    a.d_dskip, a.d_fg, a.d_dfg, a.d_z, a.d_dh_in, a.d_wb_all = ds_pair.data_ptr(), fg.data_ptr(), dfg.data_ptr(), zb.data_ptr(), dh_in.data_ptr(), wb.data_ptr()
    a.channels, a.precision = CH, 2
    a.layer, a.n_layers, a.B, a.L, a.dilation, a.in_start, a.out_start = 0, 1, B, L, d, in_start, out_start
    a.gs_out, a.ds_start, a.gz, a.gs_in = gs_out, ds_start, gz, gs_in
    native.check(lib.wn_tb_block_bwd_data(ctypes.byref(a), st), "bwd data")
    torch.cuda.synchronize()
    # reference (float64) with weights as the bf16 pairs represent them
    q = lambda w: (w.to(torch.bfloat16).double() + (w - w.to(torch.bfloat16).float()).to(torch.bfloat16).double())
    wrq, wsq, wfq, wgq = q(wr)[:, :, 0], q(ws)[:, :, 0], q(wf), q(wg)
    dsk_full = torch.zeros(B, L, CH, device=dev, dtype=torch.float64); dsk_full[:, ds_start:] = ds_q
    dz = (dh_q if have_dh else torch.zeros_like(dh_q)) @ wrq + dsk_full @ wsq              # (B, L, D): sum_r dh[r] Wr[r][c]
    fd, gd = f.double(), gg.double()
    dF, dG, z = dz * gd * (1 - fd * fd), dz * fd * gd * (1 - gd), fd * gd
    got_dfg = from_pair(dfg, 512, gz)
    print("  dF", rel(got_dfg[:, gz:, :256], dF[:, gz:]), " dG", rel(got_dfg[:, gz:, 256:], dG[:, gz:]), " z", rel(from_pair(zb, 256, gz)[:, gz:], z[:, gz:]))
    # dh_in from the quantised dFG the kernel produced
    dfg_q = got_dfg.double(); dfg_q[:, :gz] = 0
    dFq, dGq = dfg_q[:, :, :256], dfg_q[:, :, 256:]
    def shifted(x, s):                 # x[t + s], zero beyond L
        out = torch.zeros_like(x); out[:, :L - s] = x[:, s:]; return out
    dh_ref = torch.zeros(B, L, CH, device=dev, dtype=torch.float64)
    for tap in range(2):
        s = (1 - tap) * d
        dh_ref += shifted(dFq, s) @ wfq[:, :, tap] + shifted(dGq, s) @ wgq[:, :, tap]
    if have_dh:
        dh_ref[:, id_start:] += dh_q[:, id_start:]
    got_dh = from_pair(dh_in, 256, gs_in)
    print("  dh_in", rel(got_dh[:, gs_in:], dh_ref[:, gs_in:]))
    # ---- weight gradients from the kernel's own dfg / z
    zq = from_pair(zb, 256, gz).double(); zq[:, :gz] = 0
    work = torch.empty(lib.wn_tb_wgrad_workspace_bytes() // 4, device=dev)
    gws, gwr = torch.zeros(CH, CH, 1, device=dev), torch.zeros(CH, CH, 1, device=dev)
    gwf, gwg = torch.zeros(CH, CH, 2, device=dev), torch.zeros(CH, CH, 2, device=dev)
    w = native.TbWgradArgs()
    w.d_dskip, w.d_dh_out, w.d_dfg, w.d_z, w.d_h_in = ds_pair.data_ptr(), (dh_pair.data_ptr() if have_dh else None), dfg.data_ptr(), zb.data_ptr(), hin_pair.data_ptr()
    w.d_gws, w.d_gwr, w.d_gwf, w.d_gwg, w.d_work = gws.data_ptr(), gwr.data_ptr(), gwf.data_ptr(), gwg.data_ptr(), work.data_ptr()
    w.channels, w.precision = CH, 2
    w.B, w.L, w.dilation, w.in_start, w.ds_start, w.id_start, w.gz = B, L, d, in_start, ds_start, id_start, gz
    native.check(lib.wn_tb_wgrad(ctypes.byref(w), st), "wgrad")
    torch.cuda.synchronize()
    ref_ws = torch.einsum("bts,btc->sc", dsk_full, zq)
    print("  gws", rel(gws[:, :, 0], ref_ws))
    if have_dh:
        dhm = dh_q.clone(); dhm[:, :id_start] = 0
        print("  gwr", rel(gwr[:, :, 0], torch.einsum("btr,btc->rc", dhm, zq)))
    for tap in range(2):
        s = (1 - tap) * d
        hs = torch.zeros_like(hin_q); hs[:, s:] = hin_q[:, :L - s]          # h_in[t - s]
        print(f"  gwf tap{tap}", rel(gwf[:, :, tap], torch.einsum("btn,btr->nr", dFq, hs)), f" gwg tap{tap}", rel(gwg[:, :, tap], torch.einsum("btn,btr->nr", dGq, hs)))
