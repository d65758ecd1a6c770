# memcheck + racecheck over what changed last: pair density of real bases, the ultrasoft energy branch / augmentation charges,
# update_twnl, the skewed warp tiles + equal column tiles of the projector GEMMs (incl. the subspace LA that shares them)
set -x
timeout 2400 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2z_sanitize_memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_ultrasoft.py tests/test_update_twnl.py tests/test_subspace_la.py -m gpu -x -q -k "pairs or si54p or ultrasoft or update_twnl or mgo216_all_atoms or half_sphere or residual or gram or scf_iterations or fixture_device" 2>&1 | tail -5
echo "exit=$?"
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2z_sanitize_memcheck.log
tail -4 gpurun_out/r2z_sanitize_memcheck.log
timeout 2400 compute-sanitizer --tool racecheck --target-processes all --error-exitcode 9 --log-file gpurun_out/r2z_sanitize_racecheck.log python -m pytest tests/test_ultrasoft.py tests/test_update_twnl.py tests/test_gpu_parity.py -m gpu -x -q -k "ultrasoft_energy or update_twnl_vs or pairs or half_sphere_mode" 2>&1 | tail -5
echo "exit=$?"
grep -c "hazard" gpurun_out/r2z_sanitize_racecheck.log
tail -6 gpurun_out/r2z_sanitize_racecheck.log
