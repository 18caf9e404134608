"""Per-layer error report of the fused field kernel against the CPU emulation of its rounding points
(tests/helpers.emulate_field), with enough structure (worst rows / columns) to localise layout bugs.
Run on the B200 box:  python tests/diag_field_layers.py [rows]
Lives under tests/ because it uses the oracle (test infrastructure)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import nerf_oracle as O                      # noqa: E402
from tests.helpers import emulate_field, vterm_reference   # noqa: E402
import cv_nerf_b200                                       # noqa: E402

K = cv_nerf_b200.kernels
DEV = "cuda"


def packed_model(seed=0):
    coarse, _ = O.init_field_params(seed, 1.0, 5.0)
    order = ("l1", "l2", "l3", "l4", "l5", "l6", "l7", "l8", "l9", "l_alpha", "l10", "l11")
    params = []
    for n in order:
        params += [coarse[n + ".weight"].to(DEV), coarse[n + ".bias"].to(DEV)]
    return coarse, K.pack_model(params)


def layer_report(rows):
    p, packed = packed_model()
    gen = torch.Generator().manual_seed(rows)
    pts = torch.randn(rows, 3, generator=gen) * 2.
    dirs = torch.nn.functional.normalize(torch.randn(rows, 3, generator=gen), dim=-1)
    x = torch.cat([O.freq_encode(pts, 10), O.freq_encode(dirs, 4)], -1)
    vt = vterm_reference(p, dirs)
    xd = x.to(DEV).contiguous()
    vtd = K.viewdir_term(packed, x[:, 63:].contiguous().to(DEV), embedded=True)
    print(f"vterm max err {(vtd.cpu() - vt).abs().max().item():.3e}")
    ok = True
    for layer in range(9):
        raw, probe = K.mlp_fwd(packed, K.IN_EMBEDDED, xd, None, rows, 1, vtd, 1, in_stride=90, probe_layer=layer)
        torch.cuda.synchronize()
        want_raw, want = emulate_field(p, x[:, :63], vt, probe=layer)
        width = want.shape[1]
        got = probe.cpu()[:, :width]
        err = (got - want).abs()
        print(f"layer {layer}: max err {err.max().item():.3e}  mean err {err.mean().item():.3e}  "
              f"max|want| {want.abs().max().item():.3f}  nan {int(torch.isnan(got).sum())}")
        if err.max().item() > 2e-3 or torch.isnan(got).any():
            ok = False
            row_err, col_err = err.amax(1), err.amax(0)
            bad_rows = (row_err > 2e-3).nonzero().flatten()
            bad_cols = (col_err > 2e-3).nonzero().flatten()
            print(f"   bad rows {bad_rows.numel()}/{rows}: first {bad_rows[:16].tolist()}")
            print(f"   bad cols {bad_cols.numel()}/{width}: first {bad_cols[:32].tolist()}")
            r = int(row_err.argmax())
            print(f"   worst row {r}: got {got[r, :8].tolist()}")
            print(f"                want {want[r, :8].tolist()}")
            break
    raw_err = (raw.cpu() - want_raw).abs()
    print(f"raw: max err {raw_err.max().item():.3e} per-channel {raw_err.amax(0).tolist()}")
    ref = O.field_mlp(p, x)
    print(f"raw vs fp32 reference: max {(raw.cpu() - ref).abs().max().item():.3e}")
    return ok




if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    ok = layer_report(128) and layer_report(rows)
    print("LAYERS", "OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
