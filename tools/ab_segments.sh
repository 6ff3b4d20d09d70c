# A/B of the number of time slices per unit of k_dec_pair (PHE_DEC_SEGMENTS at key creation); bench lines into gpurun_out/
python tools/tail_probe.py 75776,37888,100000,37888,113664 > gpurun_out/r02_tail_order.json 2> gpurun_out/r02_tail_order.err; cat gpurun_out/r02_tail_order.json; echo
for n in 4 8 12 16; do
  PHE_DEC_SEGMENTS=$n python bench.py --no-cpu --no-secondary --no-config5 --no-api > gpurun_out/r02_seg$n.json 2> gpurun_out/r02_seg$n.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_seg$n.json')); k=d['kernels']; print('nseg $n', d['value'], d['ms_per_step'], k['k_dec_pair']['ms_total']/k['k_dec_pair']['launches'])"
done
