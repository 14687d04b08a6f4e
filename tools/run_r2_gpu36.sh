#!/bin/bash
# round-2 GPU session 36: last check of the committed tree as the driver runs it: smoke + the LIF / conv parity subset
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q -x -k "lif_tensor or binary_bit_exact or production_shape" 2>&1 | tail -1
