// kernel instances for complex length 2^7 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e7() { return build_entries<7>(); } } }
