#!/bin/bash
# round-2 GPU step 1: Fortran probe, GPU tests, C4 bench, 512-potential shard sweep over l-group splits
mkdir -p gpurun_out
(echo "== which"; which gfortran flang flang-new lfortran f2c f77 f95 g77 ifort ifx nvfortran pgfortran 2>&1; echo "== ls /usr/bin"; ls /usr/bin | grep -i -E "fortran|flang|f2c|f77|f95"; echo "== libgfortran"; find / -name "libgfortran*" -not -path "/proc/*" 2>/dev/null | head; echo "== gcc"; gcc --version | head -1; gcc -print-prog-name=f951; ls $(dirname $(gcc -print-prog-name=cc1)); echo "== nproc"; nproc; echo "== conda"; which conda mamba 2>&1; nvidia-smi -L) > gpurun_out/r2_fortran_probe.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputests1.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1_err.log
for lg in 0 1 2 4 8 16; do
  PHSH_B200_LGROUPS=$lg python bench.py --potentials 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p512_lg$lg.json 2>> gpurun_out/r2_p512_err.log
done
for mb in 2 32; do
  PHSH_B200_SLAB_MIN_MB=$mb python bench.py --potentials 512 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_p512_slab$mb.json 2>> gpurun_out/r2_p512_err.log
done
cat gpurun_out/r2_fortran_probe.txt gpurun_out/r2_gputests1.txt
for f in gpurun_out/r2_bench1.json gpurun_out/r2_p512_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" value %.4g ms %.3f e2e %.4g pageable %.4g (ratio %.3f) kernel_ms %.3f frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["pageable"]["value"], d["e2e"]["pageable"]["ratio_to_pinned"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
except Exception as e:
    print(" failed", e)
PY
done
tail -5 gpurun_out/r2_bench1_err.log gpurun_out/r2_p512_err.log
