for n in 8 16 24 32; do for s in 1; do echo "SLICES_PER_WAVE=$n slack=$s"; MSDA_B200_PACE_SLACK=$s MSDA_B200_SLICES_PER_WAVE=$n timeout 300 python scripts/time_variants.py train_b64_encoder_zeros --nodet 2>&1 | tail -1; done; done
echo "slices 16 slack 0"; MSDA_B200_PACE_SLACK=0 MSDA_B200_SLICES_PER_WAVE=16 timeout 300 python scripts/time_variants.py train_b64_encoder_zeros --nodet 2>&1 | tail -1
