# A/B of the reduced-solve variants at config 5 (and config 4): ms_total / ms_solve per optimize(10)
for m in win win2; do for cfg in "200 80000" "50 20000"; do echo "ORB_B200_LDLT=$m $cfg"; ORB_B200_LDLT=$m python scripts/lba_prof.py $cfg 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_total'], d['ms_solve'], d['solver_kind'])"; done; done
