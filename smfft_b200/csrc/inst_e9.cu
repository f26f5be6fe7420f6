// kernel instances for complex length 2^9 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e9() { return build_entries<9>(); } } }
