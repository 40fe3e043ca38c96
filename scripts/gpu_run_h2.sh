for ch in 128 52 26 13; do
  echo "== BAJES_B200_CHUNKS=$ch"
  BAJES_B200_CHUNKS=$ch python scripts/prof_calls.py C2 4096 6
  BAJES_B200_CHUNKS=$ch python scripts/prof_calls.py C2 512 6
  BAJES_B200_CHUNKS=$ch python scripts/prof_calls.py C2 64 6
done
