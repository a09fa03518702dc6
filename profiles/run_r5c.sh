set -x
mkdir -p gpurun_out
for tp in 1 2 0; do
  python profiles/prof_target.py 100000 0 decode_token_pass=$tp > gpurun_out/plain_tp$tp.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_vlc_tokens" -s 1 -c 1 -o gpurun_out/r5c_tok_tp$tp -f python profiles/prof_target.py 100000 0 decode_token_pass=$tp > gpurun_out/ncu_tp$tp.log 2>&1; echo "ncu tp$tp rc=$?"
done
