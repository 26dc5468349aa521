// host_logic_test.cpp -- the ROS-free C++ mirrors of the path's callers
// (include/rmd/keyframe_node.h, include/rmd/dataset_reader.h), host only:
// compiled with the system compiler against include/, no CUDA runtime, no GPU.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include <rmd/dataset_reader.h>
#include <rmd/keyframe_node.h>

static int failures = 0;
#define CHECK(cond) do { if(!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while(0)

struct Image { int id; };

struct ScriptedDepthmap   // records what the node asks for; convergence / distance follow a script
{
  std::vector<std::string> calls;
  std::vector<std::pair<float, float> > script;
  int i = -1;
  bool accept = true;
  float last_lambda = 0.f; int last_iters = 0;
  rmd::SE3<float> last_pose;
  bool setReferenceImage(const Image &, const rmd::SE3<float> &T, float, float) { calls.push_back("ref"); last_pose = T; return accept; }
  void update(const Image &, const rmd::SE3<float> &T) { ++i; calls.push_back("update"); last_pose = T; }
  float getConvergedPercentage() const { return script[i].first; }
  float getDistFromRef() const { return script[i].second; }
  void downloadDenoisedDepthmap(float l, int n) { calls.push_back("denoise"); last_lambda = l; last_iters = n; }
  void downloadConvergenceMap() { calls.push_back("conv"); }
};

struct CountingPublisher
{
  int clouds = 0, conv = 0;
  void publishDepthmapAndPointCloud() { ++clouds; }
  void publishConvergenceMap() { ++conv; }
};

static void test_node()
{
  ScriptedDepthmap dm;
  dm.script = {{2.f, 0.1f}, {11.f, 0.1f}, {0.f, 0.2f}, {3.f, 0.6f}, {10.f, 0.5f}};
  CountingPublisher pub;
  rmd::KeyframeNode<ScriptedDepthmap, Image, CountingPublisher> node(dm, pub, 10.0f, 0.5f, 100);
  CHECK(node.state() == rmd::State::TAKE_REFERENCE_FRAME);
  const rmd::SE3<float> T(0.9238795f, 0.f, 0.3826834f, 0.f, 1.f, 2.f, 3.f);
  for(int k = 0; k < 8; ++k) node.denseInput(Image{k}, T, 0.5f, 2.5f);
  const char *want[] = {"ref", "update", "update", "denoise", "conv", "ref", "update", "update", "denoise", "conv", "ref", "update"};
  CHECK(dm.calls.size() == 12);
  for(size_t k = 0; k < dm.calls.size() && k < 12; ++k) CHECK(dm.calls[k] == want[k]);
  CHECK(pub.clouds == 2 && pub.conv == 0 && node.state() == rmd::State::UPDATE);   // (10 %, 0.5 m) exactly: not finished
  CHECK(dm.last_lambda == 0.5f && dm.last_iters == 200);
  const rmd::SE3<float> Ti = T.inv();                                              // the Depthmap gets T_curr_world
  for(int r = 0; r < 3; ++r) for(int c = 0; c < 4; ++c) CHECK(dm.last_pose(r, c) == Ti(r, c));

  ScriptedDepthmap rejecting; rejecting.accept = false; rejecting.script.assign(20, std::make_pair(0.f, 0.f));
  CountingPublisher pub2;
  rmd::KeyframeNode<ScriptedDepthmap, Image, CountingPublisher> node2(rejecting, pub2, 10.0f, 0.5f, 3);
  for(int k = 0; k < 9; ++k) node2.denseInput(Image{k}, T, 1.f, 2.f);
  size_t refs = 0; for(const auto &c : rejecting.calls) refs += (c == "ref");
  CHECK(refs == 9 && pub2.conv == 2 && node2.numMsgs() == 1);                      // every 4th message with n = 3
}

static void test_dataset(const std::string &dir)
{
  const int w = 5, h = 3, n = 6;
  CHECK(std::system(("mkdir -p " + dir + "/images " + dir + "/depthmaps").c_str()) == 0);
  {
    std::ofstream seq((dir + "/seq.txt").c_str());
    for(int k = 0; k < n; ++k)
    {
      char name[64]; std::snprintf(name, sizeof(name), "scene_%03d.png", k);
      seq << name << " " << 0.1f * k << " " << -0.2f * k << " 1.5 0.0 0.3826834 0.0 0.9238795\n";
      char dn[64]; std::snprintf(dn, sizeof(dn), "/depthmaps/scene_%03d.depth", k);
      std::ofstream d((dir + dn).c_str());
      for(int i = 0; i < w * h; ++i) d << (100.0f + 10.0f * i + k) << " ";
    }
  }
  rmd::test::Dataset ds("seq.txt", dir);
  CHECK(ds.readDataSequence() && ds.size() == (size_t)n);
  CHECK(ds.readDataSequence(2, 5) && ds.size() == 3 && ds(0).image_file_name == "scene_002.png");
  CHECK(ds.readDataSequence(4, 0) && ds.size() == 2);
  CHECK(ds.readDataSequence(3, 2) && ds.size() == 0);
  CHECK(ds.readDataSequence());
  CHECK(ds(1).depthmap_file_name == "scene_001.depth" && ds.imagePath(ds(1)) == dir + "/images/scene_001.png");
  std::vector<float> depth;
  CHECK(ds.readDepthmap(depth, ds(3), w, h) && depth.size() == (size_t)(w * h));
  CHECK(std::fabs(depth[0] - 1.03f) < 1e-6f && std::fabs(depth[14] - 2.43f) < 1e-6f);   // centimetres -> metres
  rmd::SE3<float> T;
  ds.readCameraPose(T, ds(2));
  const float c = std::sqrt(0.5f);
  CHECK(std::fabs(T(0, 0) - c) < 1e-6f && std::fabs(T(0, 2) - c) < 1e-6f && std::fabs(T(2, 0) + c) < 1e-6f && T(1, 1) == 1.0f);
  CHECK(std::fabs(T(0, 3) - 0.2f) < 1e-6f && std::fabs(T(1, 3) + 0.4f) < 1e-6f && T(2, 3) == 1.5f);
  bool threw = false;
  try { ds(n); } catch(const std::out_of_range &) { threw = true; }
  CHECK(threw);
  CHECK(!rmd::test::Dataset().readDataSequence() && !rmd::test::Dataset("missing.txt", dir).readDataSequence());
  rmd::test::Dataset env("seq.txt");
  unsetenv("RMD_TEST_DATA_PATH");
  CHECK(!env.loadPathFromEnv());
  setenv("RMD_TEST_DATA_PATH", dir.c_str(), 1);
  CHECK(std::string(rmd::test::Dataset::getDataPathEnvVar()) == "RMD_TEST_DATA_PATH" && env.loadPathFromEnv() && env.readDataSequence());
}

int main(int argc, char **argv)
{
  test_node();
  test_dataset(argc > 1 ? argv[1] : "/tmp/rmd_host_logic_test");
  if(failures == 0) std::printf("ALL HOST LOGIC TESTS PASSED\n");
  return failures ? 1 : 0;
}
