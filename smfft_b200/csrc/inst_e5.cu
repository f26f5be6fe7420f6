// kernel instances for complex length 2^5 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e5() { return build_entries<5>(); } } }
