#!/bin/bash
bash scripts/r02_gpu21.sh
VKP_PRNG_STARTS=1 bash scripts/r02_gpu24.sh
