# Convenience targets; the authoritative recipes are hybridmc_b200/build.py (library +
# executable, nvcc sm_100a) and oracle/Makefile (CPU oracle + the reference compiled from
# /root/reference when it is present).
PYTHON ?= python

.PHONY: all lib oracle test test-gpu bench clean

all: lib oracle

lib:                       ## libhybridmc_b200.so and the hybridmc executable (in-tree)
	$(PYTHON) -m hybridmc_b200.build

oracle:                    ## test infrastructure only
	$(MAKE) -C oracle

test: lib oracle           ## CPU suite (no GPU needed)
	$(PYTHON) -m pytest tests -q -m "not gpu"

test-gpu: lib oracle       ## parity tests proper, on a B200
	$(PYTHON) -m pytest tests -q -m gpu

bench: lib                 ## one JSON line (see bench.py)
	$(PYTHON) bench.py

clean:
	rm -f hybridmc_b200/libhybridmc_b200.so hybridmc_b200/hybridmc
	$(MAKE) -C oracle clean || true
