#!/bin/bash
# Here (build container), after tools/profile_variants.sh ran on the GPU box: turn gpurun_out/<TAG>_main_*.ncu-rep into the
# tracked profiles/<TAG>_ncu_full_summary.txt and refresh profiles/dp_per_contract.json.
#   bash tools/make_profile_summary.sh r02f "header line"
TAG=$1; HEAD_LINE=${2:-"ncu --set full --clock-control none captures"}
OUT=profiles/${TAG}_ncu_full_summary.txt
declare -A N=( [greeks]=1000000 [price]=1000000 [iv]=3000000 [iv_greeks]=2000000 [analytic]=4000000 [surface]=10000000 )
{
  echo "$HEAD_LINE"
  echo "B200; bench.py workloads: greeks 1M BS, price 1M BS, iv 3M BSM cfg3, iv_greeks 2M BS, analytic 4M BS; raw reports: gpurun_out/${TAG}_main_*.ncu-rep (scratch)"
  echo
  N[iv_greeks_iv]=${N[iv_greeks]}; N[iv_greeks_greeks]=${N[iv_greeks]}; N[surface_price]=${N[surface]}; N[surface_greeks]=${N[surface]}
  for W in greeks price iv iv_greeks iv_greeks_iv iv_greeks_greeks surface_price surface_greeks analytic; do
    R=gpurun_out/${TAG}_main_$W.ncu-rep
    [ -f $R ] || continue
    python tools/ncu_summary.py $R
    python tools/ncu_opcodes.py $R ${N[$W]}
    echo "-- executed instructions per SASS function (kernel body first)"
    python tools/ncu_funcs.py $R | awk '$4 > 0'
    echo "-- call sites"
    python tools/ncu_calls.py $R 14
    echo
  done
} > $OUT
ARGS=""
for W in greeks price iv iv_greeks analytic; do R=gpurun_out/${TAG}_main_$W.ncu-rep; [ -f $R ] && ARGS="$ARGS $W=$R:${N[$W]}"; done
# the two-launch IV -> greeks route: one capture per launch, summed
R=gpurun_out/${TAG}_main_iv_greeks; [ -f ${R}_iv.ncu-rep ] && ARGS="$ARGS iv_greeks=${R}_iv.ncu-rep+${R}_greeks.ncu-rep:${N[iv_greeks]}"
R=gpurun_out/${TAG}_main_surface; [ -f ${R}_price.ncu-rep ] && ARGS="$ARGS surface=${R}_price.ncu-rep+${R}_greeks.ncu-rep:${N[surface]}"
python tools/ncu_dp_counts.py $TAG $ARGS > /dev/null
wc -l $OUT; python -c "
import json; d=json.load(open('profiles/dp_per_contract.json'))
for k,v in d['kernels'].items(): print(k, v['capture'], 'dp_inst',v['dp_inst'],'slots',v['dp_slots'],'inst',v['inst_slots'],'traffic',v['traffic_bytes'],'pipe%',round(v['fp64_pipe_pct'],1),'issue%',round(v['issue_active_pct'],1),'us',v['duration_us'])"
