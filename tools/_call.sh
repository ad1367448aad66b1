run() { PROFILE_DEVICE_OUT=1 timeout 300 python tools/profile_jpegs.py 64 gpu ${RST:-240} 3 2>&1 | tail -2 | sed 's/gpu: //' | tr '\n' ';'; echo; }
echo "--- bench picture"; run
echo "--- smooth, no noise"; PROFILE_SMOOTH=1 PROFILE_NOISE=0 run
echo "--- smooth, noise 2, q95"; PROFILE_SMOOTH=1 PROFILE_NOISE=2 PROFILE_Q=95 run
echo "--- noise 60 q98"; PROFILE_NOISE=60 PROFILE_Q=98 run
echo "--- q40"; PROFILE_Q=40 run
echo "--- 4:4:4"; PROFILE_SS=0 run
echo "--- 4:2:2"; PROFILE_SS=1 run
echo "--- grey"; PROFILE_SS=L run
echo "--- smooth no restart markers"; RST=0 PROFILE_SMOOTH=1 PROFILE_NOISE=1 run
