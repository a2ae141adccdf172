// Host-side checks of the scheduling helpers the kernels share with the launch code (no GPU needed):
// block selection by residue range (expanded_tc.cuh sel_to_blk / sel_count) -- the exact sample pass,
// the trial screen and the main screen of the screened search must partition the y blocks.
#include <cstdio>
#include <set>
#include <vector>
#include "../../raft_b200/csrc/expanded_tc.cuh"

int main()
{
  int bad = 0;
  for (int tiles_n = 0; tiles_n <= 300; ++tiles_n)
    for (int S : {2, 3, 8, 32}) {
      std::vector<std::pair<int, int>> ranges = {{0, 1}, {1, 2}, {2, S}};
      if (S == 2) ranges = {{0, 1}, {1, 2}};
      std::set<int> seen;
      for (auto [lo, hi] : ranges) {
        const int cnt = b2d::sel_count(tiles_n, S, lo, hi);
        int prev = -1;
        for (int s = 0; s < cnt; ++s) {
          const int b = b2d::sel_to_blk(s, S, lo, hi);
          if (b <= prev || b < 0 || b >= tiles_n || b % S < lo || b % S >= hi || !seen.insert(b).second) ++bad;
          prev = b;
        }
        // the next index would fall outside
        if (b2d::sel_to_blk(cnt, S, lo, hi) < tiles_n) ++bad;
      }
      if (static_cast<int>(seen.size()) != tiles_n) ++bad;
    }
  // S <= 1: identity
  for (int s = 0; s < 10; ++s) if (b2d::sel_to_blk(s, 0, 0, 0) != s || b2d::sel_count(s, 1, 0, 0) != s) ++bad;
  std::printf(bad ? "FAIL %d\n" : "PASS\n", bad);
  return bad ? 1 : 0;
}
