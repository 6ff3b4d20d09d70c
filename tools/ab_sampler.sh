for per in 0.1 5 0.1 5 0.02; do
  PHE_BENCH_CLOCK_PERIOD=$per python bench.py --no-cpu --no-secondary --no-config5 --no-api > gpurun_out/r02_samp.json 2> gpurun_out/r02_samp.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_samp.json')); k=d['kernels']; s=sum(v['ms_total']/v['launches'] for v in k.values() if isinstance(v,dict)); print('period $per', d['value'], d['ms_per_step'], 'kernels', s, 'gap', d['ms_per_step']-s, d['clocks'])"
done
