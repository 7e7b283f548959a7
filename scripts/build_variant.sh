#!/bin/bash
# Build an experimental variant of the library: scripts/build_variant.sh <name> [-DFLAG ...]
# -> build/exp/lib_<name>.so (select it at run time with ORL_LIB=...), prints registers / spills / loop span.
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/exp
out=$root/build/exp/lib_$name.so
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -prec-div=false -prec-sqrt=false -Xcompiler -fPIC -shared \
  -Xptxas -v "$@" -o $out $root/openroborl_b200/csrc/orl_env.cu 2>&1 | grep -A2 "orl_step_kernelILb0ELb0E" | grep -E "registers|spill" | tr -d "\n"; echo | sed "s/^/[$name] /"
python $root/scripts/sass_sched.py $out orl_step_kernelILb0ELb0E | grep -E "substep loop|fall-through"
