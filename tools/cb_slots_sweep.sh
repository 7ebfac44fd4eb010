python -m pytest tests -m gpu -x -q -k "chessboard or baseline_configs" 2>&1 | tail -4
for s in 64 64; do
  echo "slots $s"; EZC_CB_SLOTS=$s python bench.py --no-rig --no-cpu --models radtan5 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); h=d['headline']; print(h['chessboard_cfg1_img_s'], h['chessboard_cfg2_img_s'])"
done
echo sequential; EZC_CB_SEQUENTIAL=1 python bench.py --no-rig --no-cpu --models radtan5 --steps 3 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith(chr(123))][-1]); h=d['headline']; print(h['chessboard_cfg1_img_s'], h['chessboard_cfg2_img_s'])"
