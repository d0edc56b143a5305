O=gpurun_out; mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -25 > $O/r3_i_tests.log
cat $O/r3_i_tests.log
python tools/_single.py 20
AOG_NO_SMALL=1 python tools/_single.py 20
AOG_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r3_i_single.csv python tools/_single.py 2 > /dev/null 2>&1
python tools/summarize_ncu.py launches $O/r3_i_single.csv $O/r3_i_single.md; head -10 $O/r3_i_single.md
