#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv, prof_final.ncu-rep, bench*.json} into the tracked summaries under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1_final"
shutil.copy(os.path.join(G, "launches.csv"), os.path.join(P, f"{tag}_launches.csv"))
shutil.copy(os.path.join(G, "bench.json"), os.path.join(P, f"{tag}_bench.json"))
if os.path.exists(os.path.join(G, "bench_reference.json")):
    shutil.copy(os.path.join(G, "bench_reference.json"), os.path.join(P, f"{tag}_bench_reference.json"))
raw = subprocess.run(["ncu", "-i", os.path.join(G, "prof_final.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
d = [{w: r[hdr.index(w)] for w in want if w in hdr} for r in rows[2:]]
json.dump(d, open(os.path.join(P, f"{tag}_ncu_sweep_kernels.json"), "w"), indent=1)
f = lambda x: float(x.replace(",", ""))
lr = [r for r in csv.reader(open(os.path.join(P, f"{tag}_launches.csv"))) if len(r) > 10]
lh = lr[0]; ki = lh.index('Kernel Name'); vi = lh.index('Metric Value')
agg = collections.OrderedDict()
for r in lr[1:]:
    agg.setdefault(r[ki].split('(')[0].replace('void glrm::', '').replace('glrm::', ''), []).append(f(r[vi]))
tot = sum(sum(v) for v in agg.values())
L = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` (first 200 launches)", "",
     "`ncu --metrics gpu__time_duration.sum --clock-control none -c 200` — per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
     "| kernel | launches | total ms | share |", "|---|---|---|---|"]
L += [f"| `{k}` | {len(v)} | {sum(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% |" for k, v in agg.items()]
L += ["", "# ncu --set full of the sweep kernels of one iteration (warm-up fit)", "",
      "| kernel | grid | ms | DRAM rd+wr MB | L2->L1 GB | L2 thr % | L1 data-pipe % | DRAM thr % | fp64 pipe % | issue active % | regs | L1 hit % | L2 hit % |",
      "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for k in d:
    L.append("| `%s` | %s | %.3f | %.0f | %.2f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %.1f | %.1f |" % (
        k['Kernel Name'].replace('void ', '').replace('(SweepArgs)', ''), k['Grid Size'], f(k['gpu__time_duration.sum']),
        f(k['dram__bytes_read.sum']) + f(k['dram__bytes_write.sum']), f(k['lts__t_sectors_srcunit_tex_op_read.sum']) * 32 / 1e9,
        f(k['lts__throughput.avg.pct_of_peak_sustained_elapsed']), f(k['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']),
        f(k['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']), f(k['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']),
        f(k['smsp__issue_active.avg.pct_of_peak_sustained_active']), k['launch__registers_per_thread'], f(k['l1tex__t_sector_hit_rate.pct']), f(k['lts__t_sector_hit_rate.pct'])))
# the X sweep = kernels whose grids belong to the row side: the largest warp-tier grid marks it
xs = [k for k in d if k['Grid Size'] in ('(34174, 1, 1)', '(1799, 1, 1)') or k['Grid Size'].startswith('(34174') ]
# cluster / cta launches of the X sweep precede the 34174 launch in capture order
idx = [i for i, k in enumerate(d) if k['Grid Size'].startswith('(34174')]
xk = d[:idx[0] + 1] if idx else xs
x_dram = sum(f(k['dram__bytes_read.sum']) + f(k['dram__bytes_write.sum']) for k in xk) * 1e6
x_l2 = sum(f(k['lts__t_sectors_srcunit_tex_op_read.sum']) for k in xk) * 32
x_ms = sum(f(k['gpu__time_duration.sum']) for k in xk)
L += ["", f"X sweep = the first {len(xk)} captured launches: {x_ms:.3f} ms under ncu (serialised), DRAM traffic {x_dram/1e6:.0f} MB, "
          f"L2->L1 {x_l2/1e9:.2f} GB.",
      "", "Reading: both factors live in L2 (hit rate 93-97 %), DRAM sees only the index/value streams (a few % of DRAM peak);",
      "the sweeps are bound by the L2->SM gather path (L1 data pipe 55-75 % busy, ~9-12 TB/s out of L2), not by HBM.", "See DESIGN.md section 4.1 'Roofline'."]
open(os.path.join(P, f"{tag}_ncu_summary.md"), "w").write("\n".join(L) + "\n")
json.dump({"dram_bytes_per_step": x_dram, "l2_to_l1_bytes_per_step": x_l2,
           "what": f"dram__bytes_read.sum + dram__bytes_write.sum over the {len(xk)} launches of one X sweep, ncu --set full, profiles/{tag}_ncu_sweep_kernels.json"},
          open(os.path.join(P, "ncu_update_x_traffic.json"), "w"), indent=1)
print("\n".join(L[-12:]))
