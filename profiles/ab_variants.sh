for lib in libhybridmc_b200 libhmc_scalar libhmc_packed_u2 libhmc_scalar_u2; do
  export HMC_LIB=$PWD/hybridmc_b200/$lib.so
  echo "== $lib"
  for c in "test 8192 3" "crambin 2048 2" "50bead_9bond 2048 2"; do timeout 60 python profiles/prof_step.py $c | tail -1; done
  timeout 100 python profiles/prof_run.py test 4096 1 | tail -1
done
