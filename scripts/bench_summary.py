"""Print the numbers of one or more bench.py JSON lines that the round notes quote.  usage: bench_summary.py file.json ..."""
import json, sys
def load(p):
    for line in open(p):
        if line.startswith("{"):
            return json.loads(line)
for p in sys.argv[1:]:
    d = load(p)
    if not d:
        print(p, "no JSON line"); continue
    e = d.get("extra", {})
    pc = d.get("parity_check") or {}
    print(f"{p}: N={d['n_gpus']} cfg2 {d['ms_per_step']:.4f} ms {d['value']:.0f} GB/s frac/gpu {d['roofline']['frac']:.3f} e2e {d['e2e']['value']:.0f} "
          f"parity ok={pc.get('ok')} exact={pc.get('bit_exact_rows')}/{pc.get('rows')} err/bound={pc.get('max_err_over_bound')}")
    print("   x_exchange", d.get("x_exchange"), "step_ms", {k: round(v, 4) for k, v in (d.get("step_ms") or {}).items() if isinstance(v, float)})
    for k in ("collectives", "spmv_banded_gbs", "bicgstab_cfg3_iters_per_s", "bicgstab_cfg3_launches_per_iter", "spmv_cfg3_gbs", "spmv_cfg3_exchange",
              "bicgstab_cfg3_parity", "bicgstab_small_vs_oracle", "arnoldi_cfg4_steps_per_s", "arnoldi_cfg4_gbs", "gmres_cfg4_cycles_per_s", "gmres_cfg4_gbs",
              "spmm_cfg5_k16_ms", "spmm_cfg5_k16_gbs", "spmm_cfg5_k16_dist_ms", "spmm_cfg5_k16_dist_gbs", "spmm_cfg5_dist_error"):
        if k in e:
            print("    ", k, e[k])
