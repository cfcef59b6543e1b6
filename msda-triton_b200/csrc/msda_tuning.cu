// msda_tuning.cu -- see msda_tuning.h: the environment is read once, not on the launch path.
#include <cstdlib>
#include <mutex>

#include "msda_tuning.h"

namespace msda {

namespace {

Tuning g_tuning;
std::once_flag g_tuning_once;

int env_int(const char *name, int dflt) {
    const char *e = std::getenv(name);
    if (!e || !e[0]) return dflt;
    return std::atoi(e);
}

void read_env() {
    Tuning t;
    t.force_generic = env_int("MSDA_B200_FORCE_GENERIC", 0) != 0;
    t.slices_per_wave = env_int("MSDA_B200_SLICES_PER_WAVE", 0);
    if (t.slices_per_wave < 0) t.slices_per_wave = 0;
    t.pace_slack = env_int("MSDA_B200_PACE_SLACK", 1);
    t.wave_pacing = env_int("MSDA_B200_WAVE_PACING", 1);
    t.fwd_variant = env_int("MSDA_B200_FWD_VARIANT", -1);
    t.bwd_split = env_int("MSDA_B200_BWD_SPLIT", 0) != 0;
    t.split_slots = env_int("MSDA_B200_SPLIT_SLOTS", 0);
    t.bwd_tmem = env_int("MSDA_B200_BWD_TMEM", -1);
    t.tmem_levels = env_int("MSDA_B200_TMEM_LEVELS", 2);
    t.tmem_warps = env_int("MSDA_B200_TMEM_WARPS", 15);
    t.l1_keep_kb = env_int("MSDA_B200_L1_KEEP_KB", 100);
    t.bwd_agg = env_int("MSDA_B200_BWD_AGG", -1);
    t.bwd_dense = env_int("MSDA_B200_BWD_DENSE", -1);
    t.dense_prefetch = env_int("MSDA_B200_DENSE_PF", 2);
    t.bwd_shape = env_int("MSDA_B200_BWD_SHAPE", -1);
    t.carveout = env_int("MSDA_B200_CARVEOUT", -1);
    t.det_variant = env_int("MSDA_B200_DET_VARIANT", -1);
    g_tuning = t;
}

}  // namespace

const Tuning &tuning() {
    std::call_once(g_tuning_once, read_env);
    return g_tuning;
}

void reload_tuning() {
    std::call_once(g_tuning_once, read_env);
    read_env();
}

}  // namespace msda
