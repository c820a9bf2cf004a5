#!/usr/bin/env python3
"""Turn ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv  > profiles/rNN_launches.md
  python profiles/summarize.py kernel   gpurun_out/prof.ncu-rep  > profiles/rNN_kernel.md
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [(r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", ""))) for r in rows]
    # one step = from one k_plan_smem/k_init_bits launch sequence start to the next k_alloc (inclusive)
    # a tick ends with its last allocator kernel (k_alloc<..>, then k_alloc_groupless when the tick has one)
    ends = [i for i, (n, _) in enumerate(names) if "k_alloc" in n and (i + 1 == len(names) or "k_alloc" not in names[i + 1][0])]
    print(f"# ncu launch list ({len(names)} launches; `--metrics gpu__time_duration.sum --clock-control none`)\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    def table(seg, title):
        agg = collections.OrderedDict()
        for n, v in seg:
            a = agg.setdefault(n, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(v for _, v in agg.values())
        print(f"{title}\n")
        print("| kernel | launches | time (us) | share |\n|---|---:|---:|---:|")
        for n, (c, v) in agg.items():
            print(f"| `{n}` | {c} | {v / 1e3:.1f} | {100 * v / tot:.1f}% |")
        print(f"| **total** | {sum(c for c, _ in agg.values())} | {tot / 1e3:.1f} | 100% |\n")
    if len(ends) < 2:
        table(names, "All launches:")
        return
    segs = [names[a + 1: b + 1] for a, b in zip(ends[:-1], ends[1:])]
    # resident ticks (evg_run_resident: the bench's timed region) have no k_validate; the one-shot pipelined call
    # (evg_plan_and_alloc_batch, the e2e leg) validates and plans chunk by chunk
    resident = [s for s in segs if not any(n.startswith("k_validate") for n, _ in s)]
    chunked = [s for s in segs if any(n.startswith("k_validate") for n, _ in s)]
    if resident:
        table(resident[-1], "One resident tick (the timed region of bench.py):")
    if chunked:
        table(chunked[-1], "One chunk of the pipelined one-shot call (bench.py e2e leg; H2D/D2H copies overlap on other streams):")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
            "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct"]
    print(f"# ncu --set full: {m.get('Kernel Name', ('?', ''))[0][:90]}\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for k in keep[1:]:
        if k in m:
            print(f"| `{k}` | {m[k][0]} | {m[k][1]} |")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
    if not hi:
        return
    hdr = rows[hi[0]]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg, cur = collections.OrderedDict(), None
    for r in rows[hi[0] + 1:]:
        if len(r) <= ci:
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:95])
            agg.setdefault(cur, [0, 0, collections.Counter()])
            continue
        if cur is None or r[2] in ("-", "..."):
            continue
        try:
            ins, smp = int(r[ci]), int(r[cs])
        except ValueError:
            continue
        agg[cur][0] += ins
        agg[cur][1] += smp
        for i, h in stall_cols:
            try:
                agg[cur][2][h] += int(r[i])
            except ValueError:
                pass
    ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    print(f"\nWarp-level instructions executed: {ti}; stall samples: {ts}\n")
    print("Top source lines by stall samples:\n\n| line | samples | instr | top stalls | source |\n|---:|---:|---:|---|---|")
    for (ln, s), (ins, smp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        tops = ", ".join(f"{k[6:]} {v}" for k, v in st.most_common(2))
        print(f"| {ln} | {100 * smp / max(ts, 1):.1f}% | {100 * ins / max(ti, 1):.1f}% | {tops} | `{s.replace('|', '/')}` |")


def kernels(path):
    """Every kernel of a multi-kernel capture, one row each."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
            ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors"), ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors"),
            ("smsp__inst_executed.sum", "warp instr"), ("sm__inst_executed.avg.per_cycle_elapsed", "IPC/SM"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid")]
    ix = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none: {len(rows) - 2} launches of `{path.split('/')[-1]}`\n")
    print("| kernel | " + " | ".join(f"{t} ({units[ix[c]]})" if units[ix[c]] else t for c, t in cols if c in ix) + " |")
    print("|---|" + "---:|" * sum(c in ix for c, _ in cols))
    for r in rows[2:]:
        def fmt(v):
            try:
                x = float(v.replace(",", ""))
                return f"{x:.4g}" if abs(x) < 1e6 else f"{x:.4e}"
            except ValueError:
                return v
        print(f"| `{r[ix['Kernel Name']].split('(')[0].replace('void ', '')}` | " + " | ".join(fmt(r[ix[c]]) for c, _ in cols if c in ix) + " |")


def phases(path, units_per_launch="1"):
    """One kernel: SASS sorted by address, cut at every BAR.SYNC; per phase the share of stall samples, thread
    instructions per unit (task) and the top stall reasons.  `units_per_launch` = tasks the profiled launch processed."""
    n_units = float(units_per_launch)
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    data.sort(key=lambda r: int(r[ix["Address"]], 16))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda x: float(x) if x not in ("", "-") else 0.0
    ph, cur = [], []
    for r in data:
        cur.append(r)
        if "BAR.SYNC" in r[ix["Source"]]:
            ph.append(cur); cur = []
    if cur:
        ph.append(cur)
    tot_s = sum(num(r[ix["# Samples"]]) for r in data)
    tot_t = sum(num(r[ix["Thread Instructions Executed"]]) for r in data)
    tot_w = sum(num(r[ix["Instructions Executed"]]) for r in data)
    print(f"# {rows[0][1][:100] if len(rows[0]) > 1 else path}\n")
    print(f"Stall samples {tot_s:.0f}; warp instructions {tot_w:.4g}; thread instructions per unit {tot_t / n_units:.1f} ({n_units:.0f} units per launch).\n")
    print("Phases = stretches of SASS between consecutive `BAR.SYNC` (address order; out-of-line code such as spin loops lands in the last one).\n")
    print("| phase | SASS | samples | thread instr / unit | top stalls |\n|---:|---:|---:|---:|---|")
    for k, p in enumerate(ph):
        sm = sum(num(r[ix["# Samples"]]) for r in p)
        if sm < 0.005 * tot_s:
            continue
        ti = sum(num(r[ix["Thread Instructions Executed"]]) for r in p)
        st = collections.Counter({n: sum(num(r[ix[n]]) for r in p) for n in stalls})
        tops = ", ".join(f"{n[6:]} {100 * v / max(sm, 1):.0f}%" for n, v in st.most_common(4))
        print(f"| {k} | {len(p)} | {100 * sm / tot_s:.1f}% | {ti / n_units:.1f} | {tops} |")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "kernels": kernels, "phases": phases}[sys.argv[1]](*sys.argv[2:])
