// tests/standin/binding_standin.cpp — TEST INFRASTRUCTURE ONLY (CPU suite): the product binding bindings/delly_b200_main.cpp compiled
// together with the alignment forwarders of host_standin.cpp (the reference's own edlib / longNeedle / msa from oracle/_ref), so that
// everything the binding adds — htslib reading into record lists, option handling, the stage sequence, the htslib BCF writer — can be
// compared byte for byte with the reference's drivers on a machine without a GPU. The context entry points are served here as well
// (no device is opened). The `-m gpu` variant of the test runs the real binary (delly_b200/bin/delly_b200) on a B200.
#include "host_standin.cpp"
extern "C" {
int dgpu_ctx_create(int, dgpu_ctx** out) { *out = (dgpu_ctx*) standin_ctx(); return DGPU_OK; }
void dgpu_ctx_destroy(dgpu_ctx*) {}
const char* dgpu_last_error(dgpu_ctx*) { return "stand-in"; }
}
#include "../../bindings/delly_b200_main.cpp"
