# memcheck over the kernels added in round 2 (E_kin sums, PSDA update, Jacobi diag, v(r) producers, many-row chunked projectors)
set -x
timeout 1200 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2_sanitize.log python -m pytest tests/test_vhxc.py tests/test_subspace_la.py tests/test_gpu_parity.py tests/test_fastio.py -m gpu -x -q -k "vhxc or psda or diag or ekin or many_rows or coexist or checkpoint or scf" 2>&1 | tail -5
echo "exit=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitize.log
tail -5 gpurun_out/r2_sanitize.log
