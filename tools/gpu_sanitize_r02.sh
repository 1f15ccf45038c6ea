# compute-sanitizer over the kernels that are new or changed in round 2 (small cases): memcheck on the smoke batch, the KSA /
# CIS tests, one PM6-d fixture, the mid-size eigensolver and the large-path dimer; racecheck on the smoke batch (phase-templated
# Jacobi loop with inline ld/st.shared, bulk-copy Fock kernel) and one KSA trajectory.
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_mem_smoke.log 2>&1; echo "memcheck smoke rc=$?"; tail -3 gpurun_out/san_mem_smoke.log
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_md.py -m gpu -x -q -k "ksa_operators or ksa_scf or cis or ksa_md and methane" > gpurun_out/san_mem_ksa.log 2>&1; echo "memcheck ksa rc=$?"; tail -3 gpurun_out/san_mem_ksa.log
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pm6d_notebook_diatomics or pm6d_organics_c1" > gpurun_out/san_mem_pm6d.log 2>&1; echo "memcheck pm6d rc=$?"; tail -3 gpurun_out/san_mem_pm6d.log
timeout 1800 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_large_molecule.py -m gpu -x -q -k "mid_size or dimer or square_product and 216" > gpurun_out/san_mem_large.log 2>&1; echo "memcheck large rc=$?"; tail -3 gpurun_out/san_mem_large.log
timeout 1500 $S --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/san_race_smoke.log 2>&1; echo "racecheck smoke rc=$?"; tail -3 gpurun_out/san_race_smoke.log
timeout 1500 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_md.py -m gpu -x -q -k "ksa_md and methane" > gpurun_out/san_race_ksa.log 2>&1; echo "racecheck ksa rc=$?"; tail -3 gpurun_out/san_race_ksa.log
