"""Summarise gpurun_out/timeline_*.json (scripts/timeline.py): last replay only."""
import json, re, sys
evs = json.load(open(sys.argv[1]))
evs.sort(key=lambda x: x[2])
def nm(s):
    s = re.sub(r"\(anonymous namespace\)::", "", s)
    s = re.sub(r"^void ", "", s)
    s = re.sub(r"at::native::", "at::", s)
    m = re.match(r"([A-Za-z0-9_:]+)(<[^>]*>?)?", s)
    base = m.group(1) if m else s[:40]
    if base.startswith("gemm_bf16_tc_kernel"):
        return "gemm_bf16_tc_kernel"
    return base[:48]
# split replays at the 2 largest idle gaps (between end of everything so far and next start)
ends = 0; gaps = []
for i, e in enumerate(evs):
    if i and e[2] - ends > 0: gaps.append((e[2] - ends, i))
    ends = max(ends, e[2] + e[3])
gaps.sort(reverse=True)
cuts = sorted(i for _, i in gaps[:2])
last = evs[cuts[-1]:] if cuts else evs
t0 = last[0][2]; t1 = max(e[2] + e[3] for e in last)
busy = 0; cs = ce = None
for e in last:
    s_, e_ = e[2], e[2] + e[3]
    if cs is None: cs, ce = s_, e_
    elif s_ <= ce: ce = max(ce, e_)
    else: busy += ce - cs; cs, ce = s_, e_
busy += ce - cs
tot = {}
for e in last:
    d = tot.setdefault(nm(e[0]), [0.0, 0]); d[0] += e[3]; d[1] += 1
print("replay span %.1f us, some kernel running %.1f us (%.1f%%), %d kernels, sum of kernel time %.1f us" % (t1 - t0, busy, 100 * busy / (t1 - t0), len(last), sum(e[3] for e in last)))
for k, (d, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%9.1f us %4d x  %s" % (d, n, k))
for e in last:
    if "nccl" in e[0].lower():
        print("NCCL kernel at +%.1f us for %.1f us: %s" % (e[2] - t0, e[3], e[0][:70]))
if len(sys.argv) > 2:
    for e in last: print("%9.1f %8.1f s%-3d %s" % (e[2] - t0, e[3], e[1], nm(e[0])))
