for bits in 3072 1024; do for n in 100000 94720; do for seg in 1 16; do
  echo "bits $bits n $n seg $seg: $(PHE_DEC_SEGMENTS=$seg python tools/config5_3072.py --bits $bits --count $n 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["ms_decrypt"],2), round(d["ms_encrypt"],2))')"
done; done; done
