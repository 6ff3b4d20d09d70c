# usage: bash tools/ab_lib_full.sh <variant suffix> ... : headline, encrypt, HE add / mul and config5 per library variant
for v in "" "$@"; do
  if [ -n "$v" ]; then export PHE_B200_LIB=$PWD/pailliercryptolib_python_b200/lib/libphe_b200_$v.so; else unset PHE_B200_LIB; fi
  python bench.py --no-cpu --no-api > gpurun_out/ab_lib.json 2> gpurun_out/ab_lib.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_lib.json')); k=d['kernels']; c=d['config3']; c5=d['config5']; print('variant[$v]', round(d['value']), round(d['ms_per_step'],2), 'dec', round(k['k_dec_pair']['ms_total']/5,2), 'enc', round(k['k_encrypt_npair']['ms_total']/5,2), 'add', round(c['he_add_ops_s']/1e6,1), 'bcast', round(c['he_add_broadcast_ops_s']/1e6,1), 'mul53', round(c['he_mul53_ops_s']/1e6,3), 'mul2048', round(c['he_mul2048_ops_s']), 'c5 enc/dec', round(c5['ms_encrypt'],2), round(c5['ms_decrypt'],2), c['bit_exact_vs_oracle'])"
done
