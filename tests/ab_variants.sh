# A/B of library variants on the GPU box: bash tests/ab_variants.sh old main ... (writes gpurun_out/ab_*.json; read with
# python tests/ab_show.py old main).  A variant <v> is lib/libintopt_b200_<v>.so, built by
# INTOPT_VARIANT=<v> INTOPT_NVCC_EXTRA="-D..." bash neuripsintopt_b200/csrc/build.sh; "main" is the product library.
for v in "$@"; do
  if [ "$v" = main ]; then unset INTOPT_LIB_OVERRIDE; else export INTOPT_LIB_OVERRIDE=$PWD/neuripsintopt_b200/lib/libintopt_b200_$v.so; fi
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-alt --no-strong --regions 3 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  timeout 200 python bench_configs.py --no-dense --batch 2368 > gpurun_out/abcfg_$v.jsonl 2> gpurun_out/abcfg_$v.err
done
echo done
