// Host-side internals shared by api_dev.cc / lr_host.cc / the plugin ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include "../../../include/pdsb.h"

namespace pdsb {

void* pinned_alloc(size_t bytes);
void pinned_free(void* p);
int thread_streams(cudaStream_t* compute, cudaStream_t* copy);

// null policy of the reference (src/linear/mod.rs:34-66)
enum class NullKind { RAISE, SKIP, SKIP_WINDOW, IGNORE, FILL, FILL_WINDOW };
struct NullPolicy { NullKind kind; double fill; };
// returns 0 / 1 ("Invalid NullPolicy.")
int parse_null_policy(const char* s, NullPolicy* out);

int solver_from_string(const char* s);   // lr/mod.rs:18-27
int se_type_from_string(const char* s);  // linear_regression.rs:122-132

}  // namespace pdsb
