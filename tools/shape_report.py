"""Join the dry-run GEMM shape list of one training step with an ncu launch list (per-launch
durations) and report achieved TFLOP/s per shape class.  CPU only."""
import collections, csv, re, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pcm_b200 import ops, config, weights
from pcm_b200.step import PCMTrainStep

csv_path = sys.argv[1]
ops.DRY_RUN = []
cfg = config.SD15
sd = weights.synthetic_state_dict(cfg, 0)
st = PCMTrainStep(cfg, sd, "cpu", batch=8, height=64, width=64, multiphase=4)
ops.DRY_RUN.clear()
st.run_eager()
rec = ops.DRY_RUN
gemms = [r[1] for r in rec if r[0] == "gemm"]
print("dry-run kernels:", len(rec), "gemms:", len(gemms))
rows = list(csv.reader(open(csv_path)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
mi = hdr.index('Metric Name')
dur = []
for r in rows[hi + 1:]:
    if len(r) > vi and r[mi] == 'gpu__time_duration.sum' and \
            ('pcm_gemm_kernel' in r[ki] or 'pcm_gemm_epi2_kernel' in r[ki] or 'pcm_gemm2_kernel' in r[ki]):
        v = float(r[vi].replace(',', '')); u = r[ui]
        v *= {'us': 1e-6, 'ns': 1e-9, 'ms': 1e-3}.get(u, 1.0)
        dur.append(v)
# the ncu run contains the constructor's launches first (none are gemm) then one step
print("ncu gemm launches:", len(dur))
assert len(dur) == len(gemms)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for g, t in zip(gemms, dur):
    key = (g['M'], g['N'], g['K'], g['bn'], g['lin'])
    a = agg[key]; a[0] += 1; a[1] += t; a[2] += g.get('flop', 2.0 * g['M'] * g['N'] * g['K'])   # algorithmic (ranged K entries)
tot_t = sum(a[1] for a in agg.values()); tot_f = sum(a[2] for a in agg.values())
print(f"total {tot_t*1e3:.2f} ms, {tot_f/1e12:.2f} TFLOP, {tot_f/tot_t/1e12:.1f} TFLOP/s")
print("   ms    %time   n    M      N     K    bn lin  TF/s  tiles")
for key, (n, t, f) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    M, N, K, bn, lin = key
    tiles = ((M + 127) // 128) * ((N + bn - 1) // bn)
    print(f"{t*1e3:7.2f} {100*t/tot_t:5.1f}% {n:4d} {M:6d} {N:6d} {K:6d} {bn:4d} {lin:2d} {f/t/1e12:7.1f} {tiles:6d}")
