// b200_internal.h -- private declarations shared by the translation units of
// libb200sphinx.so.  Not installed; the public surface is include/b200sphinx.h.
#pragma once
#include "../../include/b200sphinx.h"

#include <cstdarg>
#include <cstdint>
#include <string>
#include <vector>

namespace b200 {

void set_error(const char *fmt, ...);

// sphinxbase/src/libsphinxbase/util/logmath.c restated as a value type.
struct LogMath {
    double base, log_of_base, inv_log_of_base;
    int shift;
    int width = 1;
    int32_t zero;
    std::vector<uint32_t> table;
    LogMath(double base, int shift, bool use_table);
    int32_t log(double p) const;
    int32_t ln_to_log(double ln_p) const;
    int32_t add(int32_t x, int32_t y) const;
};

}  // namespace b200
