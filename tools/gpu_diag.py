"""GPU diagnostics for the fused field kernel (run on the B200 box):
    python tools/gpu_diag.py --time                 times both field kernels on a lego-sized workload
    python tools/gpu_diag.py --ab 9,30 [--rounds N] interleaved A/B of debug-entry variants (experiments build):
                                                    milliseconds, cycles per CTA, bit-identity against the first
    python tools/gpu_diag.py --stats                wait fractions of the probe build
The per-layer error report against the CPU emulation of the kernel's rounding points lives with the tests
(python tests/diag_field_layers.py): only tests may use the oracle."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cv_nerf_b200                                       # noqa: E402
from cv_nerf_b200.model import Model                      # noqa: E402

K = cv_nerf_b200.kernels
DEV = "cuda"


def packed_model(seed=0):
    """A default-initialised field network (the reference's nn.Linear init) and its packed blob."""
    torch.manual_seed(seed)
    net = Model().to(DEV)
    return net, net.packed()


def timing():
    p, packed = packed_model()
    n_rays, S = 160000, 64
    rays = torch.zeros(n_rays, 11, device=DEV)
    rays[:, 0:3] = torch.randn(n_rays, 3, device=DEV) * .3 + torch.tensor([0., 0., 4.], device=DEV)
    rays[:, 3:6] = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV), dim=-1)
    rays[:, 6], rays[:, 7] = 2., 6.
    rays[:, 8:11] = rays[:, 3:6]
    ht = K.model_host_tail(packed)
    for S in (64, 192):
        z = K.sample_coarse(rays, S)
        vt = K.viewdir_term(packed, rays)
        ref = None
        for tag, tail in (("biases staged in shared memory", None), ("biases in the kernel parameters (production)", ht)):
            for _ in range(3):
                raw = K.mlp_fwd(packed, K.IN_RAYS, rays, z, n_rays * S, S, vt, S, host_tail=tail)
            torch.cuda.synchronize()
            if ref is None:
                ref = raw
            else:
                print(f"  max |{tag} - smem bias| =", (raw - ref).abs().max().item())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            reps = 5
            for _ in range(reps):
                K.mlp_fwd(packed, K.IN_RAYS, rays, z, n_rays * S, S, vt, S, host_tail=tail)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            flops = n_rays * S * 1186816
            print(f"mlp_fwd[{tag}] {n_rays} rays x {S} samples: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s "
                  f"({n_rays * S / ms / 1e3:.2f} M samples/s)")


def ab(variants, rounds=6):
    """Interleaved A/B timing of debug-entry variants (the power-capped chip drifts by several percent within
    a process, so variants are timed round-robin and averaged)."""
    from cv_nerf_b200 import _lib
    lib = _lib.load()
    if not _lib.has_experiments():
        sys.exit("needs the experiments build: make -C cv-nerf_b200/csrc experiments && "
                 "NERF_B200_LIB=$PWD/cv-nerf_b200/libnerf_b200_exp.so python tools/gpu_diag.py ...")
    p, packed = packed_model()
    n_rays, S = 160000, 192
    rays = torch.zeros(n_rays, 11, device=DEV)
    rays[:, 0:3] = torch.randn(n_rays, 3, device=DEV) * .3
    rays[:, 3:6] = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV), dim=-1)
    rays[:, 6], rays[:, 7] = 2., 6.
    rays[:, 8:11] = rays[:, 3:6]
    z = K.sample_coarse(rays, S)
    vt = K.viewdir_term(packed, rays)
    raw = torch.empty(n_rays * S, 4, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    tot = {v: 0. for v in variants}
    cyc = {v: 0. for v in variants}
    first = None
    stats = torch.zeros(148, 8, dtype=torch.int64, device=DEV)
    for rnd in range(rounds + 1):
        for v in variants:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                # variants >= 9 write only their total cycle count per CTA into the counter block
                rc = lib.nerf_mlp_fwd_stats(packed.data_ptr(), rays.data_ptr(), z.data_ptr(), n_rays * S, S,
                                            vt.data_ptr(), raw.data_ptr(), v, stats.data_ptr() if v >= 9 else 0, st)
                assert rc == 0, lib.nerf_b200_last_error()
            e1.record()
            torch.cuda.synchronize()
            if rnd == 0:
                if first is None:
                    first = raw.clone()
                else:
                    print(f"variant {v}: max |raw - raw(variant {variants[0]})| = {(raw - first).abs().max().item():.3e}"
                          f"  bit-identical: {bool(torch.equal(raw, first))}")
            else:
                tot[v] += e0.elapsed_time(e1) / 3
                if v >= 9:
                    rows = stats[stats[:, 5] > 0, 5].double()
                    cyc[v] += rows.mean().item() if rows.numel() else 0.
    for v in variants:
        ms = tot[v] / rounds
        print(f"variant {v}: {ms:.3f} ms  {n_rays * S * 1186816 / ms / 1e9:.0f} TFLOP/s  cycles/CTA {cyc[v] / rounds:.4e} "
              f"(mean of {rounds} interleaved rounds)")


def pipeline_stats(variants=(9, 15, 16, 14)):
    """Per-role wait-cycle breakdown of the field kernel (debug entry nerf_mlp_fwd_stats)."""
    from cv_nerf_b200 import _lib
    lib = _lib.load()
    if not _lib.has_experiments():
        sys.exit("needs the experiments build: make -C cv-nerf_b200/csrc experiments && "
                 "NERF_B200_LIB=$PWD/cv-nerf_b200/libnerf_b200_exp.so python tools/gpu_diag.py ...")
    p, packed = packed_model()
    n_rays, S = 160000, 192
    rays = torch.zeros(n_rays, 11, device=DEV)
    rays[:, 0:3] = torch.randn(n_rays, 3, device=DEV) * .3
    rays[:, 3:6] = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=DEV), dim=-1)
    rays[:, 6], rays[:, 7] = 2., 6.
    rays[:, 8:11] = rays[:, 3:6]
    z = K.sample_coarse(rays, S)
    vt = K.viewdir_term(packed, rays)
    raw = torch.empty(n_rays * S, 4, device=DEV)
    names = {1: "ring 2x32KB (production layout)", 2: "ring 1x32KB", 3: "ring 3x32KB (PE aliased, timing only)",
             4: "EXP no A-tile stores", 5: "EXP no bias loads", 6: "EXP no TMEM loads", 7: "EXP none of the three",
             9: "host tail (production inference kernel); no counters", 10: "host tail + 16-warp crew; no counters",
             13: "EXP no weight streaming + 16-warp crew; no counters", 14: "EXP no weight streaming, host tail; no counters",
             16: "host tail, earlier layout (PE tiles of their own, 2 x 32 KB ring); no counters",
             15: "production with three accumulator buffers in the epilogue; no counters", 16: "EXP no bias loads + no weight streaming; no counters",
             17: "EXP host tail, whole-warp MMA issuer; no counters", 18: "production + sampled wait profile (tools/wait_profile.py); no counters",
             11: "EXP no weight streaming (upper bound if weight slots were always ready)",
             100: "CTA pairs (cta_group::2); leader CTAs only; [6] = wait for the peer's half-chunk",
             101: "CTA pairs + 16-warp crew; leader CTAs only",
             102: "CTA pairs, tensor-map weight copies", 103: "CTA pairs + crew, tensor-map weight copies",
             200: "TS kernel (activations in TMEM); producer(empty) = MMA thread waits for its own commit (queue drain), epiX = crew wait for acc, epiY = crew hidden-epilogue busy, w_peer = wait PE"}
    st = torch.cuda.current_stream().cuda_stream
    for v in variants:
        stats = torch.zeros(148, 8, dtype=torch.int64, device=DEV)
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = lib.nerf_mlp_fwd_stats(packed.data_ptr(), rays.data_ptr(), z.data_ptr(), n_rays * S, S,
                                        vt.data_ptr(), raw.data_ptr(), v, stats.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, lib.nerf_b200_last_error()
        ms = e0.elapsed_time(e1)
        rows = stats[stats[:, 5] > 0].double()
        s = rows.mean(0).cpu() if rows.shape[0] else torch.zeros(8, dtype=torch.float64)
        tot = max(s[5].item(), 1.)
        print(f"variant {v} {names.get(v, '')}: {ms:.2f} ms  {n_rays * S * 1186816 / ms / 1e9:.0f} TFLOP/s  "
              f"clock {tot / ms / 1e3:.0f} MHz | "
              f"cycles/CTA {tot:.3e}; wait fractions: producer(empty) {s[0] / tot:.2f}  mma(a_ready) {s[1] / tot:.2f}  "
              f"mma(w_full) {s[2] / tot:.2f}  epiX(acc) {s[3] / tot:.2f}  epiY(acc) {s[4] / tot:.2f}  mma(w_peer) {s[6] / tot:.2f}  [7] {s[7] / tot:.2f}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--ab", type=str, default="", help="comma-separated debug-entry variants to time round-robin")
    ap.add_argument("--rounds", type=int, default=6)
    a = ap.parse_args()
    if a.ab:
        ab([int(x) for x in a.ab.split(",")], a.rounds)
        sys.exit(0)
    if a.stats:
        pipeline_stats()
        sys.exit(0)
    print("device:", torch.cuda.get_device_name(0), "SMs", cv_nerf_b200._lib.load().nerf_b200_sm_count())
    if a.time:
        timing()
    else:
        print("nothing to do: --time, --ab or --stats (per-layer error report: python tests/diag_field_layers.py)")
