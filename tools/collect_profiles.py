"""Copy the round-2 measurement pass (tools/final_profile.sh -> gpurun_out/r02_*) into profiles/ as tracked summaries:
JSON / text files as they are, the ncu reports as selected-metric CSVs, the decode launch list as a per-kernel table, and
profiles/r02_traffic.json (DRAM bytes per launch of the default kernel, read by bench.py)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out")
DST = os.path.join(ROOT, "profiles")

KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def copy(name, dst=None):
    s = os.path.join(SRC, name)
    if os.path.exists(s):
        shutil.copy(s, os.path.join(DST, dst or name))
        print("copied", name)
        return True
    print("missing", name)
    return False


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    return rows[0], rows[1], rows[2:]


def main():
    for f in ["r02_bench_n1.json", "r02_bench_reference_arm.json", "r02_timeline_e0.25.txt", "r02_timeline_e1.0.txt",
              "r02_ubench_acc_rate.txt", "r02_ubench_stage_cost.txt", "r02_ubench_stream_acc.txt", "r02_gemv_sweep.json",
              "r02_bench_n2_tp.json", "r02_bench_n8_tp.json", "r02_tp_tests_n2.txt", "r02_launches_decode.csv"]:
        copy(f)
    traffic = {"source": "dram__bytes_read.sum + dram__bytes_write.sum of one bucket_mul_v4_kernel launch, 4096->14336, "
                         "ncu --set full --clock-control none (profiles/r02_v4_e*_ncu_selected.csv)"}
    for e in ("0.25", "1.0"):
        rep = os.path.join(SRC, f"r02_v4_e{e}.ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        r = ncu_raw(rep)
        if not r:
            continue
        h, units, body = r
        row = body[0]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        with open(os.path.join(DST, f"r02_v4_e{e}_ncu_selected.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["metric", "value", "unit"])
            w.writerow(["Kernel Name", row[h.index("Kernel Name")]])
            for k in h:
                if k in KEEP or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"):
                    w.writerow([k, row[h.index(k)], units[h.index(k)]])
        rd = float(row[h.index("dram__bytes_read.sum")].replace(",", "")) * scale[units[h.index("dram__bytes_read.sum")]]
        wr = float(row[h.index("dram__bytes_write.sum")].replace(",", "")) * scale[units[h.index("dram__bytes_write.sum")]]
        # the raw page reports bytes in the unit of the header's second row; normalise through the units row if present
        traffic[e] = int(rd + wr)
        print("traffic", e, traffic[e])
    if len(traffic) > 1:
        json.dump(traffic, open(os.path.join(DST, "r02_traffic.json"), "w"), indent=1)
    # per-kernel table of one token
    lc = os.path.join(SRC, "r02_launches_decode.csv")
    if os.path.exists(lc):
        rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
        h = rows[0]
        ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
        seq = []
        for r in rows[1:]:
            try:
                seq.append((r[ki][:48], r[gi], float(r[vi].replace(",", ""))))
            except ValueError:
                pass
        idx = [i for i, s in enumerate(seq) if "embed_kernel" in s[0]]
        if len(idx) >= 2:
            tok = seq[idx[-2]:idx[-1]]
            agg = collections.OrderedDict()
            for k, g, v in tok:
                a = agg.setdefault((k, g), [0, 0.0])
                a[0] += 1
                a[1] += v
            tot = sum(a[1] for a in agg.values())
            with open(os.path.join(DST, "r02_decode_token_kernels.txt"), "w") as f:
                f.write("one decode token (32 layers, effort 0.25) under ncu --metrics gpu__time_duration.sum (cold, serialised):\n")
                for (k, g), a in sorted(agg.items(), key=lambda x: -x[1][1]):
                    f.write(f"{a[1]/1000:9.1f} us  {a[0]:4d} launches  {a[1]/a[0]/1000:7.2f} us each  {100*a[1]/tot:5.1f}%  {k} grid {g}\n")
                f.write(f"total {tot/1000:.1f} us in {len(tok)} launches\n")
            print(open(os.path.join(DST, "r02_decode_token_kernels.txt")).read())


if __name__ == "__main__":
    sys.exit(main())
