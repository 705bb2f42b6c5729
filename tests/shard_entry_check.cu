// Host-side check of the multi-GPU partition rule (jqc_common.cuh: shard_entry): for every (n, world) the ranks'
// entries below n cover 0..n-1 exactly once, and consecutive rounds alternate direction.  Built and run by
// tests/test_host_cpu.py::test_shard_entry_partition (nvcc, host code only).
#include <cstdio>
#include <vector>

#include "jqc_common.cuh"

int main()
{
    for (int world = 1; world <= 9; world++)
        for (int n = 0; n <= 4 * world + 3; n++) {
            std::vector<int> hits(n, 0);
            for (int rank = 0; rank < world; rank++)
                for (unsigned round = 0;; round++) {
                    const unsigned e = jqc::shard_entry(round, rank, world);
                    if (e >= (unsigned)n) {
                        // an entry past the end: every later round of this rank is past the end as well
                        if (jqc::shard_entry(round + 1, rank, world) < (unsigned)n) { std::printf("FAIL order %d %d %d\n", world, n, rank); return 1; }
                        break;
                    }
                    hits[e]++;
                }
            for (int i = 0; i < n; i++)
                if (hits[i] != 1) { std::printf("FAIL cover world=%d n=%d entry=%d hits=%d\n", world, n, i, hits[i]); return 1; }
        }
    // rank 0 takes the first entry of even rounds and the last entry of odd rounds
    if (jqc::shard_entry(0, 0, 4) != 0 || jqc::shard_entry(1, 0, 4) != 7 || jqc::shard_entry(1, 3, 4) != 4) { std::printf("FAIL direction\n"); return 1; }
    std::printf("OK\n");
    return 0;
}
