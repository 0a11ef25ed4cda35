#!/bin/bash
# One GPU visit: gpu tests, then the given bench commands; everything lands in gpurun_out/<tag>_*.
# usage: scripts/gpu_visit.sh <tag> [pytest|nopytest] -- then bench command lines on stdin, one per line: "<name> <args...>"
tag=$1; shift
mkdir -p gpurun_out
if [ "$1" = "pytest" ]; then
  python -m pytest tests -m gpu -x -q -rs > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
  tail -5 gpurun_out/${tag}_pytest.log
fi
while read -r name args; do
  [ -z "$name" ] && continue
  echo "== $name: $args"
  ( time eval "$args" ) > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.log
  echo "rc=$?"; tail -c 600 gpurun_out/${tag}_${name}.log; head -c 1500 gpurun_out/${tag}_${name}.json; echo
done
