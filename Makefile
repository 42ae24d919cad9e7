# Convenience targets (the driver uses __graft_entry__.build(), pytest and bench.py directly).
PY ?= python

build:            ## libtvd.so (nvcc, sm_100a), the oracle and, if /root/reference exists, oracle/_ref
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## oracle vs reference / goldens, host logic, gloo sharding (no GPU)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:         ## CUDA vs oracle parity (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench:            ## headline metric, one GPU
	$(PY) bench.py

results:          ## replay all 456 stored results of the reference through the CUDA path
	$(PY) scripts/replay_goldens.py

configs:          ## the five BASELINE.json configurations
	$(PY) scripts/run_configs.py --cpu

.PHONY: build test-cpu test-gpu bench results configs
