#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/s23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s23_pytest.log
tail -5 gpurun_out/s23_pytest.log
