# memcheck over the kernels added in r1l (subspace LA, current density, half-sphere projectors)
set -x
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize.log python -m pytest tests/test_subspace_la.py tests/test_current_density.py tests/test_gpu_parity.py -m gpu -x -q -k "la_vs_oracle or current_vs_oracle or gamma_half or asymmetric or energy_after" 2>&1 | tail -5
echo "exit=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/sanitize.log
tail -5 gpurun_out/sanitize.log
