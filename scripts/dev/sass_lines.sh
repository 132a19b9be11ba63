#!/bin/bash
# Static SASS instruction count per source line of kob_step_fast<6, noise> (developer tool).
set -e
cd "$(dirname "$0")/../.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xptxas -v $EXTRA -cubin -o /tmp/one_kernel.cubin scripts/dev/one_kernel.cu 2>&1 | grep -E "registers|spill|error" || true
nvdisasm -g -c /tmp/one_kernel.cubin > /tmp/one_kernel.sass
python3 scripts/dev/sass_lines.py /tmp/one_kernel.sass "$@"
