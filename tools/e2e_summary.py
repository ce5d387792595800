"""Recall@10 of the end-to-end ml-100k / SVD run (README trial-223 hyper-parameters, `--augment-training-data` = only_synthetic,
SURVEY 8c trap T1): reference (tools/ref_e2e.py, unmodified reference on the CPU) vs this repo (tools/our_e2e.py, B200).
usage: python tools/e2e_summary.py profiles/e2e_reference_ml100k.jsonl profiles/e2e_ours_ml100k_r02.jsonl > profiles/e2e_recall_r02.md"""
import json, math, sys


def load(path):
    return [json.loads(l) for l in open(path) if l.strip()]


def stats(rows, name):
    v = [r[name]["recall@10"] for r in rows]
    n = len(v)
    m = sum(v) / n
    sd = math.sqrt(sum((x - m) ** 2 for x in v) / (n - 1)) if n > 1 else float("nan")
    return n, m, sd, sd / math.sqrt(n)


ref, ours = load(sys.argv[1]), load(sys.argv[2])
print("# End-to-end Recall@10, ml-100k / SVD evaluator, README trial-223 flags, only_synthetic (same flag on both sides)\n")
print(f"reference: `{sys.argv[1]}` ({len(ref)} runs, unmodified reference on the CPU, tools/ref_e2e.py)  ")
print(f"sdrm_b200: `{sys.argv[2]}` ({len(ours)} runs on B200 with the kernels of this round, tools/our_e2e.py; the binarisation goes through K4 and is "
      "asserted equal to np.quantile + compare inside every run)\n")
print("| synthetic data | reference mean ± SE (n, sd) | sdrm_b200 mean ± SE (n, sd) | delta | SE of delta | within ±0.005 |")
print("|---|---|---|---|---|---|")
for name in ("F-SDRM", "M-SDRM", "MultiVAE++"):
    n1, m1, s1, e1 = stats(ref, name)
    n2, m2, s2, e2 = stats(ours, name)
    d, se = m2 - m1, math.sqrt(e1 * e1 + e2 * e2)
    print(f"| {name} | {m1:.4f} ± {e1:.4f} ({n1}, {s1:.4f}) | {m2:.4f} ± {e2:.4f} ({n2}, {s2:.4f}) | {d:+.4f} | {se:.4f} | {'yes' if abs(d) <= 0.005 else 'NO'} |")
t = lambda rows, k: sum(r[k] for r in rows) / len(rows)
print(f"\nWall clock per run (means): train_SDRM {t(ref, 'train_s'):.0f} s -> {t(ours, 'train_s'):.1f} s; full-resolution sampling of 843 users "
      f"{t(ref, 'sample_full_s'):.2f} s -> {1e3 * t(ours, 'sample_full_s'):.1f} ms; multi-resolution sampling {t(ref, 'sample_random_s'):.1f} s -> "
      f"{1e3 * t(ours, 'sample_random_s'):.1f} ms.")
print("\nA single run scatters by ~0.012 (sd column), so the ±0.005 tolerance of the north star is only resolvable on means of >= 20 runs per side; "
      "the reference side keeps growing in the background of a build session (CPU, ~3-5 min per run).")
