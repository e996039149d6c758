// TEST INFRASTRUCTURE (oracle/): in-memory stand-ins for the reference's
// write_snapshot / read_snapshot / read_input (src/snapshot.cc:14-131), whose
// real bodies need libhdf5.  Only used so hybridmc_test.cc's `check_snapshot`
// case links; keeps the same observable round-trip (pos, s_bias, textual
// mt19937 state, config).
#include "snapshot.h"
#include <map>
#include <sstream>

namespace {
struct Snap {
  std::vector<Vec3> pos;
  std::vector<double> s_bias;
  std::string mt;
  Config config;
};
std::map<std::string, Snap> &store() {
  static std::map<std::string, Snap> s;
  return s;
}
} // namespace

void write_snapshot(const std::string snapshot_name, const std::vector<Vec3> &pos,
                    const std::vector<double> &s_bias, Random &mt,
                    UpdateConfig &update_config) {
  std::stringstream ss;
  ss << mt;
  store()[snapshot_name] = Snap{pos, s_bias, ss.str(), update_config.config};
}

void read_snapshot(const std::string snapshot_name, std::vector<Vec3> &pos,
                   std::vector<double> &s_bias, Random &mt,
                   UpdateConfig &update_config) {
  const Snap &s = store().at(snapshot_name);
  pos = s.pos;
  s_bias = s.s_bias;
  std::stringstream ss(s.mt);
  ss >> mt;
  update_config.config = s.config;
}

void read_input(const std::string, std::vector<Vec3> &) {}
