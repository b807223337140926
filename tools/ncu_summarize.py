"""Summaries of ncu outputs for profiles/: python tools/ncu_summarize.py launches <csv> | full <ncu-rep>"""
import collections, csv, subprocess, sys

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("b200sv::", "")
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v * 1e3 if u in ("s", "second") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.3f | %.3f | %.1f %% |" % (k, c, t, t / c, 100 * t / tot))
    print("\ntotal GPU time in the listed launches: %.1f ms" % tot)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    names = [r[h.index("Kernel Name")].split("(")[0] for r in rows[2:]]
    print("| metric | " + " | ".join("launch %d" % i for i in range(len(names))) + " |\n|---|" + "---:|" * len(names))
    print("| kernel | " + " | ".join("`%s`" % n.replace("void ", "").replace("b200sv::", "") for n in names) + " |")
    for i, name in enumerate(h):
        if name in WANT or ("issue_stalled" in name and name.endswith("per_issue_active.ratio")):
            vals = [r[i] for r in rows[2:]]
            if all(v in ("0", "0.000000") for v in vals):
                continue
            print("| `%s` (%s) | %s |" % (name, rows[1][i], " | ".join(vals)))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, kern, mix = None, None, {}
    for r in rows:
        if r and r[0] == "Kernel Name":
            kern = r[1].split("(")[0]
            mix.setdefault(kern, collections.Counter())
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if kern and hdr and len(r) == len(hdr):
            s = r[1].strip()
            op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
            mix[kern][op] += int(r[hdr.index("Instructions Executed")])
    for k, c in mix.items():
        tot = sum(c.values())
        print("\nExecuted warp-instruction mix, `%s`: " % k.replace("void ", "").replace("b200sv::", "") +
              ", ".join("%s %.1f %%" % (op, 100 * n / tot) for op, n in c.most_common(8)))

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
