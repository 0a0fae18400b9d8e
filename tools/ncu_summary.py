#!/usr/bin/env python
"""Turn the ncu outputs of a gpurun call into the small text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<round>_launches.txt
    python tools/ncu_summary.py raw      gpurun_out/prof.ncu-rep  > profiles/<round>_<kernel>_raw.txt
    python tools/ncu_summary.py stalls   gpurun_out/prof.ncu-rep  > profiles/<round>_<kernel>_stalls.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max",
            "sm__pipe_tensor_cycles_active.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__pipe_fma_cycles_active.avg.pct",
            "sm__pipe_alu_cycles_active.avg.pct", "smsp__issue_active.avg.pct", "sm__warps_active.avg.per_cycle_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
            "lts__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct", "dram__throughput.avg.pct",
            "gpu__dram_throughput.avg.pct", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "smsp__inst_executed.sum",
            "sm__throughput.avg.pct", "l1tex__t_bytes.sum", "smsp__sass_thread_inst_executed_op_ffma", "sm__inst_executed_pipe_tensor"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(ix):
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(r[ix["Metric Unit"]], 1.0)
        a = agg.setdefault(name, [0, 0.0, r[ix["Grid Size"]], r[ix["Block Size"]]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms "
          "(cold-cache, serialised: compare SHARES)")
    print(f"{'launches':>8} {'total ms':>10} {'avg us':>9} {'share':>7}  grid block  kernel")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{a[0]:8d} {a[1]:10.3f} {1e3 * a[1] / a[0]:9.1f} {100 * a[1] / tot:6.1f}%  {a[2]} {a[3]}  {k}")


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none, raw page of {rep}; one column per captured launch")
    print("kernel:", [r[ix["Kernel Name"]][:60] for r in rows[2:]])
    for i, h in enumerate(hdr):
        if any(h.startswith(k) for k in RAW_KEYS):
            print(f"{h} [{units[i]}] = {[r[i] for r in rows[2:]]}")


def stalls(rep, top=40):
    rows = ncu_csv(rep, "source", ("--print-source", "sass"))
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    hdr = rows[secs[0] + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    end = secs[1] if len(secs) > 1 else len(rows)
    data = [r for r in rows[secs[0] + 2:end] if len(r) > 10]
    names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    agg = sorted(((s, sum(int(r[ix[s]]) for r in data)) for s in names), key=lambda x: -x[1])
    print(f"# {rows[secs[0]][1]}: {tot} warp samples, {len(data)} SASS instructions")
    print("stall reasons:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in agg[:8]))
    ops = collections.Counter()
    for r in data:
        ops[r[1].strip().split()[1 if r[1].strip().startswith("@") else 0].split(".")[0]] += int(r[ix["Instructions Executed"]])
    print("warp-instructions executed by opcode:", ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))
    print(f"{'samples':>8} {'executed':>10}  instruction  [top stall reasons]")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
        st = sorted(((s[6:], int(r[ix[s]])) for s in names if int(r[ix[s]]) > 0), key=lambda x: -x[1])[:3]
        print(f"{r[ix['# Samples']]:>8} {r[ix['Instructions Executed']]:>10}  {r[1].strip()[:72]:72s} {st}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "stalls": stalls}[sys.argv[1]](sys.argv[2])
