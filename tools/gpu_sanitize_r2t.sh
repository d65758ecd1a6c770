# memcheck + racecheck (shared-memory hazards) over the tensor-memory kernels (k_plane_t incl. its opt-in variants, k_plane_f, k_zcol_*_t,
# k_ycols_t), the reordered GEMM stages and the ultrasoft entry points
set -x
timeout 2400 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2t_sanitize_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_ultrasoft.py tests/test_subspace_la.py -m gpu -x -q -k "mgo216_compiled or mgo216_all_atoms or mgo216_shape_vs_oracle or si54p or au992 or ultrasoft or residual or kpoint_cubic" 2>&1 | tail -5
echo "exit=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2t_sanitize_memcheck.log
tail -4 gpurun_out/r2t_sanitize_memcheck.log
timeout 2400 compute-sanitizer --tool racecheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2t_sanitize_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mgo216_compiled or si54p_shape_gamma_real[True]" 2>&1 | tail -5
echo "exit=$?"
grep -c "hazard" gpurun_out/r2t_sanitize_racecheck.log
tail -6 gpurun_out/r2t_sanitize_racecheck.log
