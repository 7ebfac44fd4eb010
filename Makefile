# Top-level convenience targets.  The driver's entry points are __graft_entry__.build() / smoke() and bench.py.
PY ?= python
SAN ?= /usr/local/cuda/bin/compute-sanitizer

.PHONY: build test test-gpu sanitize bench

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:
	$(PY) -m pytest tests -x -q -m gpu

# compute-sanitizer over the small fixtures (needs a GPU): memcheck, racecheck (shared-memory hazards in the merge /
# union-find / reduction kernels) and synccheck.  Each tool's summary goes to profiles/sanitize_<tool>.txt.
sanitize:
	@mkdir -p profiles
	for tool in memcheck racecheck synccheck; do \
	  $(SAN) --tool $$tool --error-exitcode 9 --print-limit 20 $(PY) tools/sanitize_fixtures.py > profiles/sanitize_$$tool.txt 2>&1; \
	  echo "$$tool exit $$?" >> profiles/sanitize_$$tool.txt; tail -4 profiles/sanitize_$$tool.txt; \
	done

bench:
	$(PY) bench.py
