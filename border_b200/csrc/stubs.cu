// stubs.cu -- agents whose device path is not built yet fail loudly at create time.
#include "agent.cuh"
namespace bb {
Agent* make_sac(const bb_sac_cfg&) { throw Error("SAC device path not built yet"); }
Agent* make_iqn(const bb_iqn_cfg&) { throw Error("IQN device path not built yet"); }
}
