"""Compact table from a tools/ncu_summary.py 'kernels' file: one row per captured launch.
usage: python tools/ncu_table.py <summary.txt> [label1,label2,...] > profiles/xxx.txt"""
import re
import sys

src = sys.argv[1]
labels = sys.argv[2].split(",") if len(sys.argv) > 2 else []
blocks, cur = [], {}
for line in open(src):
    if line.startswith("-" * 20):
        if cur:
            blocks.append(cur)
        cur = {}
        continue
    if line.startswith("#"):
        continue
    m = re.match(r"(\S+)\s+(.*)", line.rstrip())
    if m:
        cur[m.group(1) if m.group(1) != "Kernel" else "Kernel Name"] = m.group(2)
if cur:
    blocks.append(cur)


def num(b, k):
    v = b.get(k, "0").split()
    try:
        x = float(v[0].replace(",", ""))
    except Exception:
        return 0.0, ""
    return x, (v[1] if len(v) > 1 else "")


def mb(b, k):
    x, u = num(b, k)
    return x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)


def us(b):
    x, u = num(b, "gpu__time_duration.sum")
    return x * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)


print(f"# from {src}: ncu --set full --clock-control none (one row per captured launch; times under ncu are serialised and")
print("# cold-cache: compare shares and counters, absolute times come from bench.py / the probe tools)")
print(f"{'#':>3} {'kernel':44s} {'us':>8} {'DRAM rd MB':>10} {'wr MB':>8} {'DRAM %pk':>8} {'tensor %':>8} {'issue %':>8} {'regs':>5} {'grid':>6} {'smem KB':>8}  stalls per issue (long_sb / barrier / wait / short_sb)")
for i, b in enumerate(blocks):
    name = b.get("Kernel Name", b.get("Name", "?"))
    name = re.sub(r"^Name\s+", "", name)
    name = re.sub(r"void |unnamed>::|\(.*", "", name)[:44]
    lab = f"  [{labels[i]}]" if i < len(labels) else ""
    st = "/".join(f"{num(b, 'smsp__average_warps_issue_stalled_' + k + '_per_issue_active.ratio')[0]:.2f}"
                  for k in ("long_scoreboard", "barrier", "wait", "short_scoreboard"))
    print(f"{i:3d} {name:44s} {us(b):8.1f} {mb(b, 'dram__bytes_read.sum'):10.1f} {mb(b, 'dram__bytes_write.sum'):8.1f} "
          f"{num(b, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[0]:8.1f} "
          f"{num(b, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[0]:8.1f} "
          f"{num(b, 'smsp__issue_active.avg.pct_of_peak_sustained_active')[0]:8.1f} "
          f"{int(num(b, 'launch__registers_per_thread')[0]):5d} {int(num(b, 'launch__grid_size')[0]):6d} "
          f"{num(b, 'launch__shared_mem_per_block_dynamic')[0]:8.1f}  {st}{lab}")
