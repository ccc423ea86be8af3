// Writes a small file with h5mini.hpp for the reader test (tests/test_host_layer.py): h5mini_selftest <out.h5> [group]
#include "h5mini.hpp"
int main(int argc, char **argv) {
    if (argc < 2) return 1;
    h5mini::Writer w(argc > 2 ? argv[2] : "");
    std::vector<double> part(3 * 4 * 7), swarm(3 * 16), pred(2 * 1 * 6);
    for (size_t k = 0; k < part.size(); k++) part[k] = 0.5 * (double)k - 3.25;
    for (size_t k = 0; k < swarm.size(); k++) swarm[k] = 1.0 / (1.0 + (double)k);
    for (size_t k = 0; k < pred.size(); k++) pred[k] = -(double)k;
    w.add("part", {3, 4, 7}, part);
    w.add("swarm", {3, 16}, swarm);
    w.add("pred", {2, 1, 6}, pred);
    w.add("swarm_fishNet", {2, 4}, std::vector<double>(8, 2.5));
    return w.write(argv[1]) ? 0 : 2;
}
