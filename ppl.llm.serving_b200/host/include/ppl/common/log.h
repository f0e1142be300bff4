// LOG(level) << ... stream macro of ppl.common (EXTERNAL).  One line per statement, written to stderr
// under a mutex; level filtered by the environment variable PPL_LOG_LEVEL (DEBUG/INFO/WARNING/ERROR,
// default INFO).
#ifndef B2LLM_SHIM_PPL_COMMON_LOG_H_
#define B2LLM_SHIM_PPL_COMMON_LOG_H_

#include "retcode.h"

#include <sstream>
#include <stdint.h>
#include <string>

namespace ppl { namespace common {

enum {
    LOG_LEVEL_DEBUG = 0,
    LOG_LEVEL_INFO = 1,
    LOG_LEVEL_WARNING = 2,
    LOG_LEVEL_ERROR = 3,
    LOG_LEVEL_FATAL = 4,
};

int GetCurrentLogLevel();
void SetCurrentLogLevel(int);

class LogMessage final {
public:
    LogMessage(int level, const char* file, int line);
    ~LogMessage(); // emits the line
    template <typename T>
    LogMessage& operator<<(const T& v) {
        if (enabled_) {
            os_ << v;
        }
        return *this;
    }
    LogMessage& operator<<(std::ostream& (*manip)(std::ostream&)) {
        if (enabled_) {
            os_ << manip;
        }
        return *this;
    }

private:
    bool enabled_;
    int level_;
    std::ostringstream os_;
};

}} // namespace ppl::common

#define LOG(level) ::ppl::common::LogMessage(::ppl::common::LOG_LEVEL_##level, __FILE__, __LINE__)

#endif
