"""Turn the measured-parity log of a GPU test run (tests/helpers.record -> $NERF_B200_PARITY_LOG, JSON lines)
into the two text tables kept under profiles/:  python tools/parity_report.py parity.jsonl profiles/r02"""
import collections
import json
import sys


def main():
    rows = [json.loads(l) for l in open(sys.argv[1])]
    prefix = sys.argv[2]
    by = collections.defaultdict(list)
    for r in rows:
        by[r["kind"]].append(r)
    with open(prefix + "_frame_parity.txt", "w") as fh:
        w = lambda *a: print(*a, file=fh)
        w("# Measured on B200 by `pytest -m gpu` (tests/test_gpu_parity_configs.py, tests/test_gpu_render.py).")
        w("# GPU = the whole frame through render(c2w=...); reference = the CPU oracle (pinned bit-exactly to the real")
        w("# reference) on a random pixel subset, same uniform draws.  flip = rays whose far-sample density changes sign")
        w("# between the BF16 and the fp32 forward (SURVEY.md App. C); tolerance: max-abs rgb <= 1e-2 (north_star).")
        w()
        w(f"{'case':18s} {'frame':9s} {'output':8s} {'checked':>7s} {'max|err|':>10s} {'excl.flips':>10s} {'>1e-2':>6s} {'flips':>5s} "
          f"{'PSNR dB':>8s} {'raw rel-L2':>10s} {'max|raw ref|':>12s}")
        for r in by["frame_parity"]:
            w(f"{r['case']:18s} {r['frame']:9s} {r['output']:8s} {r['checked']:7d} {r['max_all']:10.2e} {r['max_noflip']:10.2e} "
              f"{r['n_gt_1e-2']:6d} {r.get('n_flip', 0):5d} {r['psnr_vs_oracle']:8.1f} {r['raw_rel_l2']:10.2e} {r['raw_absmax_ref']:12.4g}")
        w()
        w("PSNR delta against a synthetic target (north_star: <= 0.1 dB):")
        for r in by["frame_psnr_delta"]:
            w(f"  {r['case']:18s} {r['delta_db']:.2e} dB")
        w()
        w("Fixtures recorded from the REAL reference (tests/golden/make_golden*.py), rays of the fixture through render(rays=...):")
        w(f"{'fixture':18s} {'output':8s} {'rays':>5s} {'max|err|':>10s} {'excl.flips':>10s} {'>1e-2':>6s} {'flips':>5s}")
        for r in by["render_fixture"]:
            w(f"{r['fixture']:18s} {r['output']:8s} {r['rays']:5d} {r['max_all']:10.2e} {r['max_noflip']:10.2e} {r['n_gt_1e-2']:6d} "
              f"{r.get('n_flip', 0):5d}")
        for r in by["sharpened_weights"]:
            w()
            w("Sharpened weights (300 train steps on a 0/1 checkerboard): " + json.dumps(r))
        for r in by["frame_vs_bf16_emulation"]:
            w("  vs the CPU emulation of BF16 tensor-core math: " + json.dumps(r))
        for r in by["multirank"]:
            w()
            w("2-rank run (tests/test_gpu_multirank.py): " + json.dumps(r.get("rank0", r)))
    with open(prefix + "_grad_parity.txt", "w") as fh:
        w = lambda *a: print(*a, file=fh)
        w("# Gradient parity measured on B200 by `pytest -m gpu` (tests/test_gpu_backward.py).")
        w("# emu  = against the fp64 restatement of the kernels' own arithmetic (BF16 operands, masks from the saved activations)")
        w("# fp32 = against the fp32 autograd reference; fp32-masks = the BF16 arithmetic with the fp32 forward's ReLU masks")
        w("#        against the fp32 reference (what is left is rounding alone: the rest of `fp32` is ReLU-mask flips)")
        w()
        masks = {(r["rows"], r["tensor"]): r for r in by["grad_stage_fp32_masks"]}
        w(f"{'rows':>5s} {'tensor':16s} {'emu rel-L2':>10s} {'emu cos':>10s} {'fp32 rel-L2':>11s} {'fp32 cos':>9s} {'fp32-masks rel-L2':>17s}")
        for r in by["grad_stage"]:
            m = masks.get((r["rows"], r["tensor"]), {})
            w(f"{r['rows']:5d} {r['tensor']:16s} {r['emu_rel_l2']:10.2e} {r['emu_cos']:10.7f} {r['fp32_rel_l2']:11.4f} {r['fp32_cos']:9.5f} "
              f"{m.get('rel_l2', float('nan')):17.4f}")
        w()
        w("Fraction of ReLU masks that differ between the BF16 and the fp32 forward (per layer):")
        for r in by["relu_mask_flips"]:
            w("  " + json.dumps(r))
        w()
        w("End to end (render -> loss -> backward through the drop-in surface vs the reference's autograd, 96 / 64 rays):")
        w(f"{'fixture':12s} {'net':7s} {'tensor':16s} {'norm rel err':>12s} {'rel-L2':>8s} {'cos':>9s}")
        for r in by["grad_e2e"]:
            w(f"{r['fixture']:12s} {r['net']:7s} {r['tensor']:16s} {r['norm_rel_err']:12.4f} {r['rel_l2']:8.4f} {r['cos']:9.5f}")
        for r in by["grad_e2e_whole"]:
            w("whole gradient vector: " + json.dumps(r))
        worst = collections.defaultdict(float)
        for r in by["grad_partial_tiles"]:
            worst[r["rows"]] = max(worst[r["rows"]], r["rel_l2"])
        w()
        w("Partial tiles (1 / 127 / 129 rows), worst rel-L2 over the 24 tensors: " + json.dumps(worst))


if __name__ == "__main__":
    main()
