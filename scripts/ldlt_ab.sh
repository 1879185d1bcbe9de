for f in 0 1 3 4 5; do echo "flags $f"; ORB_B200_LDLT_FLAGS=$f python scripts/lba_prof.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_total'], d['ms_solve'])"; done
for f in 0 1; do echo "prof flags $f"; ORB_B200_LDLT_FLAGS=$f ORB_B200_LDLT_PROF=1 python scripts/lba_prof.py 2>&1 | tail -4 | head -2; done
