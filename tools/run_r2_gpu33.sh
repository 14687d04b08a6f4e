#!/bin/bash
# round-2 GPU session 33: the LIF_TENSOR parity tests with the added T = 1 shape
timeout 600 python -m pytest tests -m gpu -q -x -k "lif_tensor" 2>&1 | tail -4 | cut -c1-250
