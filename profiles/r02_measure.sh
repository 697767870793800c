#!/usr/bin/env bash
# Round-2 measurements, as run on the GPU box:  gpurun --timeout 1500 -- 'bash profiles/r02_measure.sh all'
# Text outputs only (gpurun_out/ is capped at 64 MiB): the .ncu-rep files stay in /tmp on the box and are summarised
# there by profiles/summarize_rep.py.  Nothing printed under ncu is used as a bench value.
# Parts: tests | bench | sweeps | ncu | ncudense | all
set -u
PART=${1:-all}
O=gpurun_out
T="timeout"
mkdir -p "$O" /tmp/rep
want() { [ "$PART" = all ] || [ "$PART" = "$1" ]; }

if want tests; then
  $T 900 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"
  $T 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_smoke.txt 2>&1; echo "smoke rc=$?"
fi
if want bench; then
  $T 600 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
  $T 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "ref rc=$?"
fi
if want sweeps; then
  $T 400 python bench_configs.py --batch 2368 --thr 0.1 > $O/r02_configs_thr0.1.jsonl 2> $O/r02_configs_thr0.1.err; echo "cfg rc=$?"
  $T 600 python bench_configs.py --batch 1024 --thr 0 > $O/r02_configs_thr0.jsonl 2> $O/r02_configs_thr0.err; echo "cfg0 rc=$?"
  $T 600 python bench_dense.py --shapes 256x256,256x1024,256x4096,512x256,512x1024,1024x256,1024x1024,2048x128,2048x512 --parity 0 \
      > $O/r02_dense_sweep.jsonl 2> $O/r02_dense_sweep.err; echo "dense rc=$?"
  $T 600 python bench_dense.py --shapes 256x1024,1024x256 --thr 0 --parity 1 > $O/r02_dense_sweep_thr0.jsonl 2> $O/r02_dense_sweep_thr0.err; echo "dense0 rc=$?"
  $T 300 python bench_learner.py > $O/r02_learner_steps.jsonl 2> $O/r02_learner_steps.err; echo "learner rc=$?"
  $T 300 python tests/gpu_energy_ladder.py 0 1 2 3 > $O/r02_energy_ladder.txt 2>&1; echo "ladder rc=$?"
  $T 300 python tests/gpu_converged_parity.py > $O/r02_converged_parity.txt 2>&1; echo "conv rc=$?"
  $T 300 python tests/gpu_small_batch_probe.py > $O/r02_small_batch_probe.jsonl 2> $O/r02_small_batch_probe.err; echo "small rc=$?"
fi
if [ "$PART" = ncudense ]; then
  export INTOPT_DENSE_GRAPH=0
  DENSE="python bench_dense.py --shapes 1024x256 --steps 2 --parity 0"
  $T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/rep/dense1024_launches.csv $DENSE > /dev/null 2>&1
  python profiles/launch_summary.py /tmp/rep/dense1024_launches.csv > $O/r02_dense1024_launches_summary.txt 2>&1
  cap() {
    local name=$1 k=$2 skip=$3; shift 3
    $T 900 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/$name -k "regex:$k" -s $skip -c 1 "$@" > /tmp/rep/$name.log 2>&1
    python profiles/summarize_rep.py /tmp/rep/$name.ncu-rep $O/r02_${name}_ncu.txt $name > /dev/null 2>&1 || echo "summary of $name failed"
  }
  cap dense_syrk 'dense_syrk' 3 $DENSE   # (launches 1, 2 of an iteration are the masked retry rungs: 3 is a full one)
  cap dense_solve 'dense_solve_kernel' 8 $DENSE
fi
if want ncu; then
  BENCH="python bench.py --steps 2 --warmup 1 --regions 1 --no-cpu --no-alt --no-strong --no-parity"
  $T 300 $BENCH > /dev/null 2>&1 || echo "bench (no ncu) failed"
  $T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $BENCH > /dev/null 2>&1
  python profiles/launch_summary.py $O/r02_launches.csv > $O/r02_launches_summary.txt 2>&1
  # (ncu does not see inside the conditional while-graph of the dense forward: host-driven loop for the captures)
  export INTOPT_DENSE_GRAPH=0
  DENSE="python bench_dense.py --shapes 1024x256 --steps 2 --parity 0"
  $T 300 $DENSE > /dev/null 2>&1 || echo "dense (no ncu) failed"
  $T 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/rep/dense1024_launches.csv $DENSE > /dev/null 2>&1
  python profiles/launch_summary.py /tmp/rep/dense1024_launches.csv > $O/r02_dense1024_launches_summary.txt 2>&1
  cap() {  # name, kernel regex, launches to skip, command...
    local name=$1 k=$2 skip=$3; shift 3
    $T 900 ncu --set full --clock-control none --import-source on -f -o /tmp/rep/$name -k "regex:$k" -s $skip -c 1 "$@" > /tmp/rep/$name.log 2>&1
    python profiles/summarize_rep.py /tmp/rep/$name.ncu-rep $O/r02_${name}_ncu.txt $name > /dev/null 2>&1 || echo "summary of $name failed"
  }
  cap energy_fwd 'intopt_forward.*kernel' 2 $BENCH
  cap energy_bwd 'intopt_backward.*kernel' 2 $BENCH
  cap twitter_fwd 'intopt_forward.*kernel' 1 python tests/gpu_twitter_step.py
  cap dense_syrk 'dense_syrk' 3 $DENSE   # (launches 1, 2 of an iteration are the masked retry rungs: 3 is a full one)
  cap dense_cholcol 'dense_chol_col_kernel' 8 $DENSE   # (block column 8 of the first factorisation)
  cap dense_solve 'dense_solve_kernel' 8 $DENSE
fi
ls -la $O | head -60
du -sh $O
