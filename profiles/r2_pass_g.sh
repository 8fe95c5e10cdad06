#!/bin/bash
O=gpurun_out/${1:-r2g}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bamdec.py -m gpu -x -q > $O/pytest_bamdec.log 2>&1; echo "pytest rc=$?" >> $O/pytest_bamdec.log
tail -40 $O/pytest_bamdec.log
