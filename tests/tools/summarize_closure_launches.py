"""Summarise ONE closure (the launches between two Adam updates) out of an `ncu --metrics gpu__time_duration.sum --csv`
launch list of bench.py.  usage: python tests/tools/summarize_closure_launches.py launches.csv [out.md]"""
import collections, csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors='replace') if l.startswith('"'))]
hdr = rows[0]
ki, vi, mi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('Metric Unit')
seq = []
for r in rows[1:]:
    if r[mi] != 'gpu__time_duration.sum':
        continue
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(u, 1)
    seq.append((r[ki], ns))
adam = [i for i, (k, _) in enumerate(seq) if 'multi_tensor_apply' in k]
# groups of consecutive Adam kernels
groups, cur = [], []
for i in adam:
    if cur and i != cur[-1] + 1:
        groups.append(cur); cur = []
    cur.append(i)
if cur:
    groups.append(cur)
if len(groups) >= 2:
    a, b = groups[-2][-1] + 1, groups[-1][0]
    closure = seq[a:b]
else:
    # `bench.py --ncu-closure` under `ncu --profile-from-start off`: exactly one closure + its optimizer update
    closure = seq


def klass(k):
    if k.startswith('ast::') or 'ast::' in k.split('(')[0]:
        return 'ours: ' + re.sub(r'<.*', '', k.split('ast::')[1].split('(')[0])
    if 'cutlass' in k or 'xmma' in k or 'convolve' in k or 'cudnn' in k:
        return 'cuDNN conv (fprop / dgrad)'
    if 'nchwToNhwc' in k or 'nhwcToNchw' in k or 'nhwcAddPadding' in k or 'convertTensor' in k:
        return 'cuDNN layout / padding helpers'
    if 'nccl' in k.lower():
        return 'NCCL'
    return 'torch: ' + re.sub(r'<.*', '', k)[:40]


agg = collections.OrderedDict()
for k, ns in closure:
    c = klass(k)
    agg.setdefault(c, [0, 0.0])
    agg[c][0] += 1; agg[c][1] += ns
tot = sum(v[1] for v in agg.values())
lines = [f'One closure = {len(closure)} launches, {tot / 1e6:.2f} ms summed device time (cold-cache, serialised: compare shares).', '',
         '| class | launches | ms | share |', '|---|---:|---:|---:|']
for c, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'| {c} | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f} % |')
ours = sum(v[1] for c, v in agg.items() if c.startswith('ours'))
lines += ['', f"This library's kernels: {ours / 1e6:.2f} ms = {100 * ours / tot:.1f} % of the closure's device time."]
text = '\n'.join(lines)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(text + '\n')
