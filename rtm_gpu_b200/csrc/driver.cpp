// Drop-in driver (placeholder until the full file surface lands in this round).
#include "../../include/rtm_b200.h"
int rtm_fail(int code, const char* fmt, ...);
extern "C" int rtm_run_driver(const char* run_file, int ngpu, int batch, int verbose)
{
    (void)run_file; (void)ngpu; (void)batch; (void)verbose;
    return rtm_fail(RTM_ERR_STATE, "rtm_run_driver: not built yet");
}
