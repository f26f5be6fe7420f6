// kernel instances for complex length 2^6 (one translation unit per size: parallel build)
#include "registry.hpp"
namespace smfft { namespace host { EntryList entries_e6() { return build_entries<6>(); } } }
