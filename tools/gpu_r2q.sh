#!/usr/bin/env bash
# Session Q (1 GPU): the doubt-bitmask variant of the fast-lookup kernel: int-mode parity tests, then rates at N = 2^20 and 131072.
set -uo pipefail
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "int or lut or quant or dropin or explain or parity" > $O/int_tests.log 2>&1; echo "pytest rc=$?"; tail -n 5 $O/int_tests.log
timeout 300 python tools/time_modes.py 1048576 int8_sim,int4_sim > $O/int_rates.log 2>&1; echo "rates rc=$?"; cat $O/int_rates.log
timeout 300 python tools/time_modes.py 131072 int8_sim,int4_sim >> $O/int_rates.log 2>&1; tail -n 4 $O/int_rates.log
echo "--- previous build (8-entry queue)" >> $O/int_rates.log
NB_B200_LIB=tools/variants/libnb_queue.so timeout 300 python tools/time_modes.py 1048576 int8_sim,int4_sim >> $O/int_rates.log 2>&1
NB_B200_LIB=tools/variants/libnb_queue.so timeout 300 python tools/time_modes.py 131072 int8_sim,int4_sim >> $O/int_rates.log 2>&1
echo "--- new build again" >> $O/int_rates.log
timeout 300 python tools/time_modes.py 1048576 int8_sim,int4_sim >> $O/int_rates.log 2>&1
cat $O/int_rates.log
