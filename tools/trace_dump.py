"""Phase timeline of CTA 0 of the fused kernel (EVA_SM100_TRACE=1): per-phase cycle counts averaged over items."""
import ctypes
import os
import sys
from collections import defaultdict

os.environ['EVA_SM100_TRACE'] = '1'
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from efficient_attention import _abi  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device('cuda', 0)
layer = bench.build_layer(dev, torch.float16)
x = torch.randn(B, 28, 28, 192, device=dev, dtype=torch.float16)
with torch.no_grad():
    q, k, v, _ = layer._qkv_heads(x.reshape(B, 784, 192))
    geom = _abi.eva_geometry(q, seq_shape=(28, 28), window=7, ext=0, chunk=4, chunk_ext=0)
    ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()
    for _ in range(3):
        out, path = _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)
    torch.cuda.synchronize()
assert path == 1
lib = _abi.load()
names = {1: 'item start', 2: 'pool ready', 3: 'means written', 4: 'linear ready', 5: 'omega/kbar written', 20: 'beta ready',
         21: 'beta tile written', 101: 'mma item start', 102: 'pool issued', 103: 'linear issued', 104: 'omega waited', 120: 'stats waited'}
names.update({10: 'pass2: |k|^2 of all rows done', 11: 'pass2: phi-logits ready', 12: 'pass2: logits written, P2 cleared',
              13: 'pass2: exchange barrier', 14: 'pass2: P2 tiles written', 110: 'phi-logit MMAs issued', 111: 'beta MMAs issued'})
names.update({270: 'e1: prev store drained', 271: 'e1: barrier 1', 272: 'e1: rows staged', 273: 'e1: barrier 2'})
names.update({240: 'LN: tmem loaded', 241: 'LN: math done', 242: 'LN: k side written', 243: 'LN: barrier passed', 250: 'p1: S loaded', 251: 'p1: max done', 252: 'p1: exp/pack done', 253: 'p1: O loaded'})
for pr in range(8):
    names[30 + 4 * pr] = f'pair {pr} S ready'
    names[31 + 4 * pr] = f'pair {pr} P written'
    names[32 + 4 * pr] = f'pair {pr} O ready'
    names[33 + 4 * pr] = f'pair {pr} stored'
    names[130 + 2 * pr] = f'pair {pr} S issued'
    names[131 + 2 * pr] = f'pair {pr} PV issued'
for which, label in ((0, 'compute thread 0'), (1, 'MMA thread')):
    buf = (ctypes.c_ulonglong * 8192)()
    rc = lib.eva_debug_read_trace(buf, which, 8192)
    n = int(buf[8191])
    ev = [(int(buf[i]), int(buf[i + 1])) for i in range(0, n, 2)]
    print(f'== {label}: {len(ev)} events, rc {rc}')
    # split per item
    start_ev = 1 if which == 0 else 101
    items, cur = [], []
    for e in ev:
        if e[0] == start_ev and cur:
            items.append(cur)
            cur = []
        cur.append(e)
    if cur:
        items.append(cur)
    print(f'   items traced: {len(items)}; total cycles first->last: {ev[-1][1] - ev[0][1]}')
    dur = defaultdict(list)
    for it in items[1:]:   # skip the first (cold) item
        for (e0, t0), (e1, t1) in zip(it[:-1], it[1:]):
            dur[(e0, e1)].append(t1 - t0)
    full = [it[-1][1] - it[0][1] for it in items[1:]]
    if full:
        print(f'   cycles per item (start -> last event): mean {sum(full) / len(full):.0f}')
    order = []
    for it in items[1:2]:
        order = [(a[0], b[0]) for a, b in zip(it[:-1], it[1:])]
    for key in order:
        d = dur[key]
        print(f'   {str(names.get(key[0], key[0])):>24s} -> {str(names.get(key[1], key[1])):<24s} {sum(d) / len(d):9.0f} cyc  (n={len(d)})')
    if len(items) > 2:
        gaps = [b[0][1] - a[-1][1] for a, b in zip(items[1:-1], items[2:])]
        print(f'   gap last event -> next item start: mean {sum(gaps) / len(gaps):.0f} cyc')

# ---- load latency per ring position: TMA issue (slot acquired) -> MMA thread saw the tile --------------------
def read(which):
    buf = (ctypes.c_ulonglong * 8192)()
    lib.eva_debug_read_trace(buf, which, 8192)
    n = int(buf[8191])
    return [(int(buf[i]), int(buf[i + 1])) for i in range(0, n, 2)]
tma, mma = read(2), read(1)
def split(ev, start_ev):
    items, cur = [], []
    for e in ev:
        if e[0] == start_ev and cur:
            items.append(cur)
            cur = []
        cur.append(e)
    if cur:
        items.append(cur)
    return items
ti, mi = split(tma, 1000), split(mma, 101)
k = min(len(ti), len(mi)) - 1
print(f'== loads: per ring position, issue time relative to the first issue of the item (T = TMA warp, M = MMA warp), issue -> seen by the MMA thread')
rows = {}
for j in range(1, k):
    t0 = ti[j][0][1]
    issued = {e - 1000: (t, 'T') for e, t in ti[j] if 1000 <= e < 2000}
    issued.update({e - 1000: (t, 'M') for e, t in mi[j] if 1000 <= e < 2000})
    seen = {e - 2000: t for e, t in mi[j] if 2000 <= e < 3000}
    for pos, (t, who) in issued.items():
        rows.setdefault(pos, []).append((t - t0, (seen[pos] - t) if pos in seen else None, who))
for pos in sorted(rows):
    r = rows[pos]
    lat = [x[1] for x in r if x[1] is not None]
    print(f'   pos {pos:3d} {r[0][2]}: issued at +{sum(x[0] for x in r) / len(r):8.0f}   issue->seen ' + (f'{sum(lat) / len(lat):8.0f}' if lat else '       -'))
