// kernel instances for complex length 2^11 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e11() { return build_entries<11>(); } } }
