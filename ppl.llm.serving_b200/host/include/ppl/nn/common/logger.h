// ppl/nn/common/logger.h (EXTERNAL): the reference includes it for LOG().
#ifndef B2LLM_SHIM_PPL_NN_COMMON_LOGGER_H_
#define B2LLM_SHIM_PPL_NN_COMMON_LOGGER_H_
#include "ppl/common/log.h"
#include "ppl/common/retcode.h"
#include <stdint.h>
#endif
