"""profiles/r01_ncu_ffn_summary.json (bench.py's roofline.traffic source) from a whole-step `ncu --set full` summary.

    python tools/make_ffn_summary.py gpurun_out/ncu_step_v21_summary.jsonl profiles/r01_ncu_step_v21_raw.csv 65536 17 4

The step capture averages the launches of one kernel name over the step: T-1 full launches (nodes*S rows) and the
dead-row-eliminated last timestep (nodes rows), so `rows` is that average and traffic / rows is per row."""
import json
import sys


def main():
    src, raw_name, nodes, S, T = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rows = ((T - 1) * nodes * S + nodes) / T
    out = {}
    for line in open(src):
        r = json.loads(line)
        for key in ("ffn_tc_wgrad_kernel", "ffn_tc_dgrad_kernel", "ffn_tc_fwd_kernel"):
            if key in r["kernel"]:
                assert r["launches"] == T, (key, r["launches"])
                out[key] = {"rows": rows, "duration_ms": r["duration_us"] / 1e3, "dram_read_GB": r["dram_read_MB"] / 1e3,
                            "dram_write_GB": r["dram_write_MB"] / 1e3, "traffic_bytes_per_launch": r["traffic_bytes_per_launch"],
                            "sm__pipe_tensor_cycles_active_pct": r["tensor_pct"], "gpu__dram_throughput_pct": r["dram_pct"],
                            "registers_per_thread": r["regs"], "launches_averaged": r["launches"],
                            "source": "%s (ncu --set full of one train step, bench.py --nodes %d: %d launches per kernel averaged, "
                                      "%.0f rows per launch on average)" % (raw_name, nodes, T, rows)}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
