# memcheck + racecheck (shared-memory hazards) over the tensor-memory kernels (k_plane_t, k_zcol_*_t, k_ycols_t)
set -x
timeout 1500 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2t_sanitize_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216_compiled or mgo216_all_atoms or si54p or au992" 2>&1 | tail -5
echo "exit=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2t_sanitize_memcheck.log
tail -4 gpurun_out/r2t_sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2t_sanitize_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216_compiled" 2>&1 | tail -5
echo "exit=$?"
grep -c "hazard" gpurun_out/r2t_sanitize_racecheck.log
tail -6 gpurun_out/r2t_sanitize_racecheck.log
