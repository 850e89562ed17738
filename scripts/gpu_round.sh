#!/bin/bash
# One GPU-box visit: GPU tests, the bench line, the training bench line, and the ncu evidence of the lookahead kernels.
# Usage (under gpurun): bash scripts/gpu_round.sh <tag> [skip_tests]
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
  tail -5 $out/${tag}_pytest.log
fi
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 $out/${tag}_bench.err
python bench.py --workload train --steps 5 --warmup 3 > $out/${tag}_train.json 2> $out/${tag}_train.err; echo "train rc=$?"
tail -c 600 $out/${tag}_train.err; cat $out/${tag}_train.json
python scripts/prof_step.py 4 > $out/${tag}_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_ -s 9 -c 3 -o $out/${tag}_prof python scripts/prof_step.py 4 > $out/${tag}_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $out/${tag}_ncu.log
