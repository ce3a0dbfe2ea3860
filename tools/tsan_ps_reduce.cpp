// ThreadSanitizer / ASAN driver for the one-pass parameter-server reducer (tools/sanitize.sh tsan).
// Raw files stand in for safetensors (the reducer only sees byte offsets): 3 workers, 2 tensors, data at
// an odd byte offset, two rounds so that the momentum / cumulative inputs are exercised as well.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "core/ps_reduce.hpp"

static void write_file(const std::string& path, uint64_t lead, const std::vector<float>& v) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) { std::perror(path.c_str()); std::exit(2); }
  std::vector<char> pad(lead, ' ');
  std::fwrite(pad.data(), 1, pad.size(), f);
  std::fwrite(v.data(), sizeof(float), v.size(), f);
  std::fclose(f);
}
static std::vector<float> read_file(const std::string& path, uint64_t lead, uint64_t n) {
  std::vector<float> v(n);
  FILE* f = std::fopen(path.c_str(), "rb");
  std::fseek(f, static_cast<long>(lead), SEEK_SET);
  if (std::fread(v.data(), sizeof(float), n, f) != n) std::exit(3);
  std::fclose(f);
  return v;
}

int main() {
  const char* tmp = std::getenv("TMPDIR");
  const std::string dir = std::string(tmp ? tmp : "/tmp") + "/hypha-tsan-ps-" + std::to_string(::getpid());
  if (::mkdir(dir.c_str(), 0700) != 0) { std::perror("mkdir"); return 2; }
  const uint64_t n0 = (1u << 20) + 37, n1 = 1000, lead = 13, N = 3;
  const std::vector<uint64_t> numels{n0, n1}, offs{lead, lead + 4 * n0};
  const double lr = 0.7, mu = 0.9;
  std::vector<float> m_ref(n0 + n1, 0.f), c_ref(n0 + n1, 0.f);
  int bad = 0;
  std::string prev_m, prev_c;
  for (int rnd = 0; rnd < 2; ++rnd) {
    std::vector<std::string> paths;
    std::vector<std::vector<float>> data(N, std::vector<float>(n0 + n1));
    for (uint64_t k = 0; k < N; ++k) {
      for (uint64_t i = 0; i < n0 + n1; ++i) data[k][i] = std::sin(0.001f * static_cast<float>(i) + static_cast<float>(k + 3 * rnd));
      paths.push_back(dir + "/w" + std::to_string(k) + "-" + std::to_string(rnd));
      write_file(paths.back(), lead, data[k]);
    }
    const std::string r = std::to_string(rnd);
    const std::string ou = dir + "/u" + r, om = dir + "/m" + r, oc = dir + "/c" + r;
    for (auto& p : {ou, om, oc}) write_file(p, lead, std::vector<float>(n0 + n1, -1.f));
    auto st = hypha::ps_fused_round(paths, std::vector<std::vector<uint64_t>>(N, offs), numels, prev_m,
                                    prev_m.empty() ? std::vector<uint64_t>{} : offs, prev_c,
                                    prev_c.empty() ? std::vector<uint64_t>{} : offs, ou, om, oc, offs, lr, mu, 8);
    auto u = read_file(ou, lead, n0 + n1), m = read_file(om, lead, n0 + n1), c = read_file(oc, lead, n0 + n1);
    double worst = 0;
    for (uint64_t i = 0; i < n0 + n1; ++i) {
      const float g = (data[0][i] + data[1][i] + data[2][i]) * (1.0f / 3.0f);
      const float mm = rnd == 0 ? g : static_cast<float>(mu) * m_ref[i] + g;
      const float uu = static_cast<float>(lr) * (static_cast<float>(mu) * mm + g);
      m_ref[i] = mm;
      c_ref[i] += uu;
      worst = std::fmax(worst, std::fabs(u[i] - uu) + std::fabs(m[i] - mm) + std::fabs(c[i] - c_ref[i]));
    }
    std::printf("ps round %d: %llu elements, %d threads, max err %.3g -> %s\n", rnd, (unsigned long long)st.elements,
                st.threads, worst, worst < 1e-5 ? "ok" : "FAIL");
    bad += !(worst < 1e-5) || st.elements != n0 + n1;
    prev_m = om;
    prev_c = oc;
  }
  const std::string cmd = "rm -rf '" + dir + "'";
  if (std::system(cmd.c_str()) != 0) return 4;
  return bad;
}
