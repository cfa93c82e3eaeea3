#!/bin/bash
# A/B sweep of the 8-bit-field ADC scan variants (lanes per document x entries per 8-bit accumulation) on one GPU
out=${1:-gpurun_out/tune_u8.txt}
: > $out
echo "== 16-bit fields (adc_scan_cf_kernel)" >> $out
RC_ADC_FIELDS=16 QB_M=48,32,64,96 python tools/quick_bench.py adc >> $out 2>&1
for cfg in "48 4 2" "48 4 3" "48 4 6" "48 8 2" "48 8 3" "48 8 6" "32 4 2" "32 4 4" "32 8 2" "32 8 4" \
           "64 4 2" "64 4 4" "64 8 2" "64 8 4" "96 4 2" "96 4 3" "96 8 2" "96 8 3"; do
  set -- $cfg
  echo "== M=$1 lpd=$2 acc=$3" >> $out
  RC_ADC_U8_LPD=$2 RC_ADC_U8_ACC=$3 QB_M=$1 python tools/quick_bench.py adc 2>&1 | grep "^adc\|rror" >> $out
done
