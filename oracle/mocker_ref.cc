// mocker_ref.cc -- TEST / BENCH INFRASTRUCTURE: the synthetic producer for the CPU arms.
//
// bench.py --impl reference (and the cpu_baseline leg) must not map the product library, yet they need the very same
// byte stream the GPU arm parses.  The stream is a pure function of (config, record index) defined once, in
// flow-pipeline_b200/csrc/mocker_gen.h (it follows mocker/mocker.go:57-102 field for field); this file instantiates
// that header as a tiny host-only shared library.  It is an input generator, not part of the algorithm under test:
// nothing here decodes or aggregates.
#include <stddef.h>
#include <stdint.h>

#include "../flow-pipeline_b200/csrc/mocker_gen.h"

extern "C" int fo_mocker_host(const fa_mocker_config *cfg, uint64_t first, uint32_t n, uint8_t *buf, size_t cap, uint32_t *offsets,
                              size_t *bytes)
{
    if (!cfg || !bytes) return -1;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += fa_mocker_record_len(*cfg, first + i);
    *bytes = (size_t)total;
    if (total > 0xFFFFFFF0ull) return -1;
    if (total > cap || !buf || !offsets) return -4;
    uint8_t *p = buf;
    for (uint32_t i = 0; i < n; i++) {
        offsets[i] = (uint32_t)(p - buf);
        p = fa_mocker_record_put(*cfg, first + i, p);
    }
    offsets[n] = (uint32_t)(p - buf);
    return 0;
}
