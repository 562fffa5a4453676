"""Render the -DATTN_TRACE clock trace of attn_bwd_tc_kernel (CTA 0; a diagnosis tool).  usage: attn_trace_view.py LOG [n_events]"""
import collections
import sys

lines = [l.split() for l in open(sys.argv[1]) if l.startswith("TR")]
n_show = int(sys.argv[2]) if len(sys.argv) > 2 else 90
launches, cur, prev = [], None, None
for _, w, i, c in lines:
    w, i, c = int(w), int(i), int(c)
    if w == 0 and prev != 0:
        cur = {0: [], 1: [], 2: []}
        launches.append(cur)
    cur[w].append((i, c))
    prev = w
L = launches[-1]
t0 = min(c for w in L for _, c in L[w])
names = {0: {0: "h0 wait sdp", 2: "h0 got sdp", 4: "h0 computed", 6: "h0 got bar_c", 8: "h0 arrived pds", 10: "h0 epi start", 11: "h0 acc in regs", 13: "h0 got bar_c(epi)", 12: "h0 epi done", 14: "h0 item start", 15: "h0 got aux", 16: "h0 dQ in regs", 17: "h0 got stage_free"},
         1: {20: "L loads done", 21: "L got bar_item", 22: "L arrived aux", 0: "h1 wait sdp", 2: "h1 got sdp", 4: "h1 computed", 6: "h1 got bar_c", 8: "h1 arrived pds", 10: "h1 epi start", 11: "h1 acc in regs", 13: "h1 got bar_c(epi)", 12: "h1 epi done"},
         2: {1: "P wait pds0", 3: "P got pds0", 4: "P issued C", 6: "P issued A(next)", 5: "P issued dQ"}}
def name(w, i):
    # compute warps (roles 0 = warp 0, 1 = warp 8) walk both halves of an iteration: events of half 1 carry id + 32
    if w < 2 and i not in (20, 21, 22):
        base = names[0].get(i & 31, str(i & 31))
        return base.replace("h0", "h1") if i >= 32 else base
    return names[w].get(i, str(i))


ev = sorted(((c - t0) & 0xFFFFFFFF, w, i) for w in L for i, c in L[w])
for t, w, i in ev[:n_show]:
    print(f"{t:8d}  {'                    ' * w}{name(w, i)}")
print("span", ev[-1][0])
for w in L:
    seq = [(i, (c - t0) & 0xFFFFFFFF) for i, c in L[w]]
    ph = collections.defaultdict(list)
    for (i1, c1), (i2, c2) in zip(seq, seq[1:]):
        ph[(i1, i2)].append(c2 - c1)
    print("role", w)
    for k, v in sorted(ph.items()):
        print("  ", k, "n", len(v), "avg", sum(v) // len(v), "min", min(v), "max", max(v))
