# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests -q -m gpu

bench: build
	$(PY) bench.py

clean:
	$(MAKE) -C traceit_b200/csrc clean
	$(MAKE) -C traceit_b200/host clean
	$(MAKE) -C oracle clean

.PHONY: build test-cpu test-gpu bench clean
