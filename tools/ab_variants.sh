for lib in "" opengate_b200/csrc/variant_t1024_s2048.so opengate_b200/csrc/variant_t256_s3072.so opengate_b200/csrc/variant_t512_s2048.so; do
  echo "== lib=$lib"
  for c in c3 c1; do env ${lib:+GATE_B200_LIB=$lib} timeout 120 python tools/profile_run.py --config $c --histories 20000000 --launches 2 | cut -c1-20,100-400; done
done
