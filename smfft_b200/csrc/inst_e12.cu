// kernel instances for complex length 2^12 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e12() { return build_entries<12>(); } } }
