python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 4
python -m pytest tests/test_gpu_bf16_chain.py -q 2>&1 | tail -n 2
