// absl::string_view stand-in (abseil is absent in this image): the reference's tokenizer headers only need the
// type to exist (tokenizer_impl_sp.h:50).
#ifndef B2LLM_SHIM_ABSL_STRING_VIEW_H_
#define B2LLM_SHIM_ABSL_STRING_VIEW_H_
#include <string_view>
namespace absl {
using string_view = std::string_view;
}
#endif
