# decrypt rate against batch size, whole units (PHE_DEC_SEGMENTS=1) vs time-sliced units (default) -> gpurun_out/r02_sweep_*.json
S=19000,25000,31000,37888,44000,50000,56832,63000,70000,75776,82000,90000,100000,107000,113664,120000,131072
PHE_DEC_SEGMENTS=1 python tools/tail_probe.py $S > gpurun_out/r02_sweep_whole.json 2> gpurun_out/r02_sweep_whole.err
python tools/tail_probe.py $S > gpurun_out/r02_sweep_sliced.json 2> gpurun_out/r02_sweep_sliced.err
PHE_DEC_SEGMENTS=4 python tools/tail_probe.py $S > gpurun_out/r02_sweep_sliced4.json 2> gpurun_out/r02_sweep_sliced4.err
PHE_DEC_SEGMENTS=16 python tools/tail_probe.py $S > gpurun_out/r02_sweep_sliced16.json 2> gpurun_out/r02_sweep_sliced16.err
python - <<'PY'
import json
a=json.load(open('gpurun_out/r02_sweep_whole.json')); b=json.load(open('gpurun_out/r02_sweep_sliced.json')); c=json.load(open('gpurun_out/r02_sweep_sliced16.json')); d=json.load(open('gpurun_out/r02_sweep_sliced4.json'))
for k in a: print(k, round(a[k]['waves_per_launch']*2,2), round(a[k]['k_dec_pair_ms'],2), round(b[k]['k_dec_pair_ms'],2), round(c[k]['k_dec_pair_ms'],2), round(d[k]['k_dec_pair_ms'],2))
PY
