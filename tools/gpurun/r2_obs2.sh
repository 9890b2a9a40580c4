mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -k "running_observables or vortex_configurations_of" > gpurun_out/r2_pytest_obs2.txt 2>&1
grep -v "^$" gpurun_out/r2_pytest_obs2.txt | tail -60 | cut -c1-220
