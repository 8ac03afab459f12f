python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do
TOD_TC_RELU=0 TOD_RELU_TAB=0 python tools/ab_step.py 64 300
python tools/ab_step.py 64 300
done
TOD_TC_PLAN=1 python tools/ab_step.py 64 1 2>&1 | grep TC_PLAN | awk '{print $NF, $0}' | grep -o "mode=[0-9]*" | sort | uniq -c
