set -x
mkdir -p gpurun_out
# ncu --set full of the encoder in its new default form (encode_rounds=4) and with the factorised transform (=2), 65 536 frames (one launch)
for f in 4 2; do
  python profiles/prof_target.py 65536 0 encode_rounds=$f > gpurun_out/plain_c$f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_encode16v2" -s 1 -c 1 -o gpurun_out/r6c_enc_form$f -f python profiles/prof_target.py 65536 0 encode_rounds=$f > gpurun_out/ncu_full_c$f.log 2>&1; echo "ncu form $f rc=$?"
done
ls -la gpurun_out | grep r6c
