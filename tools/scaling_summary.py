"""profiles/r2_scaling.txt from the committed bench lines profiles/r2_bench_n{1,2,4,8}.json:
    python tools/scaling_summary.py > profiles/r2_scaling.txt
Efficiency here is value_N / (N * value_1) of the committed lines; the driver computes its own from its own runs."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    print(f"# Round 2, final build: `bench.py --gpus N --steps 20 --warmup 5` on the SAME 10 M trajectories "
          f"(T=8, d=l=4, n_c=8, 20 % missing)")
    print("# columns: N | ms per EM iteration | evaluations/s | efficiency v_N/(N v_1) | roofline.frac (algorithmic "
          "flops / measured DFMA peak) | e2e ms per step | H2D GB/s per rank (min) | per-fit ms (first call of the "
          "process) | C1 s | C2 ms | C4 ms | C5 s | replay on one GPU: labels equal, max rel param err | class gate | "
          "M-step statistics of the timed iteration | the same iteration with every cluster re-summed, ms\n")
    v1 = None
    for N in (1, 2, 4, 8):
        path = os.path.join(ROOT, "profiles", f"{tag}_bench_n{N}.json")
        if not os.path.exists(path):
            continue
        d = json.load(open(path))
        v = d["value"]
        v1 = v1 or v
        c, e = d["configs"], d["e2e"]
        pf = e["per_fit"]
        rp = d["parity"].get("multi_gpu")
        cg = d["parity"].get("class_sharded")
        print(f"{N} | {d['ms_per_step']:.3f} | {v:.3e} | {v / (N * v1):.3f} | {d['roofline']['frac']:.3f} | "
              f"{e['ms_per_step']:.1f} | {e['h2d_probe']['gbs_per_rank_min']:.1f} | {pf['ms_total']:.1f} "
              f"({pf['ms_total_first_call']:.1f}) | {c['C1']['seconds']:.3f} | {c['C2']['ms_per_step']:.3f} | "
              f"{c['C4']['ms']:.1f} | {c['C5']['seconds_per_sweep_iteration']:.3f} | "
              + (f"{rp['assignments_equal']}, {rp['params_max_rel_err']:.1e}" if rp else "-")
              + f" | {cg['ok'] if cg else '-'}"
              + f" | {d['statistics']['mode_of_timed_iteration']} | {d['statistics']['ms_per_step_full_statistics']:.3f}")
        st, wc, ck = d["state"], d["workload_check"], d["clocks"]
        print(f"    state: sha256(labels)={str(st.get('assign_sha256'))[:16]} Q={st.get('Q'):.6f} "
              f"min_count={wc.get('min_count')} status={wc.get('status')} invalid={d.get('invalid')} "
              f"clocks={ck.get('sm_mhz')}/{ck.get('sm_max_mhz')} {ck.get('reasons')}")
    for N in (1, 8):
        path = os.path.join(ROOT, "profiles", f"{tag}_bench_reference_arm_n{N}.json")
        if os.path.exists(path):
            d = json.load(open(path))
            print(f"\nreference arm (CPU port, bounded sample) on the N={N} box: {d['value']:.3e} evals/s on "
                  f"{d['cpu_baseline']['cores']} cores")


if __name__ == "__main__":
    main()
