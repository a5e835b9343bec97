set -x
mkdir -p gpurun_out
python tools/step_profile.py gpurun_out/step_profile_c2_v6.txt c2 > /dev/null 2>&1
python tools/step_profile.py gpurun_out/step_profile_c3_v4.txt c3 > /dev/null 2>&1
head -1 gpurun_out/step_profile_c2_v6.txt gpurun_out/step_profile_c3_v4.txt
