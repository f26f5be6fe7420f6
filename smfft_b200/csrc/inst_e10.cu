// kernel instances for complex length 2^10 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e10() { return build_entries<10>(); } } }
