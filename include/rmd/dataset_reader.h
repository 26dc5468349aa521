// rmd/dataset_reader.h -- rmd::test::Dataset (test/dataset.h:33-87,
// test/dataset.cpp:62-213) without boost::filesystem, Eigen and OpenCV: the
// on-disk format of the REMODE data sets.
//   <dataset path>/<sequence file>    one line per frame: `<image file> tx ty tz qx qy qz qw`
//   <dataset path>/images/<image file>        8-bit gray image (decoded by the caller's image library)
//   <dataset path>/depthmaps/<stem>.depth     ground-truth depth in centimetres, text, row-major
#ifndef RMD_DATASET_READER_H
#define RMD_DATASET_READER_H

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <rmd/se3.cuh>

namespace rmd
{
namespace test
{

struct DatasetEntry            // test/dataset.h:33-54
{
  std::string image_file_name, depthmap_file_name;
  float translation[3];        // x y z
  float quaternion[4];         // x y z w, the file's order
  DatasetEntry() : translation{0.f, 0.f, 0.f}, quaternion{0.f, 0.f, 0.f, 1.f} {}
};

class Dataset
{
public:
  explicit Dataset(const std::string &sequence_file = std::string(), const std::string &dataset_path = std::string())
    : dataset_path_(dataset_path), sequence_file_(sequence_file) {}

  static const char *getDataPathEnvVar() { return "RMD_TEST_DATA_PATH"; }   // test/dataset.h:84

  bool loadPathFromEnv()                                                       // test/dataset.cpp:199-208
  {
    const char *p = std::getenv(getDataPathEnvVar());
    if(!p) return false;
    dataset_path_ = p;
    return true;
  }

  // Lines [start, end) of the sequence file; end == 0 means "to the end" (test/dataset.cpp:83-138).
  bool readDataSequence(size_t start = 0, size_t end = 0)
  {
    if(dataset_path_.empty() || sequence_file_.empty()) return false;
    dataset_.clear();
    std::ifstream f(join(dataset_path_, sequence_file_).c_str());
    if(!f.is_open()) return false;
    std::string line;
    for(size_t line_cnt = 0; std::getline(f, line); ++line_cnt)
    {
      if(line_cnt < start || !(line_cnt < end || end == 0)) continue;
      std::stringstream ls(line);
      DatasetEntry e;
      ls >> e.image_file_name;
      e.depthmap_file_name = e.image_file_name.substr(0, e.image_file_name.find('.') + 1) + "depth";   // :106
      ls >> e.translation[0] >> e.translation[1] >> e.translation[2];
      ls >> e.quaternion[0] >> e.quaternion[1] >> e.quaternion[2] >> e.quaternion[3];
      dataset_.push_back(e);
    }
    return true;
  }

  std::string imagePath(const DatasetEntry &e) const { return join(join(dataset_path_, "images"), e.image_file_name); }

  void readCameraPose(SE3<float> &pose, const DatasetEntry &e) const         // :156-167
  {
    pose = SE3<float>(e.quaternion[3], e.quaternion[0], e.quaternion[1], e.quaternion[2],
                      e.translation[0], e.translation[1], e.translation[2]);
  }

  // width*height floats, metres (the file holds centimetres: z / 100.0f, :185)
  bool readDepthmap(std::vector<float> &depthmap, const DatasetEntry &e, size_t width, size_t height) const
  {
    std::ifstream f(join(join(dataset_path_, "depthmaps"), e.depthmap_file_name).c_str());
    if(!f.is_open()) return false;
    depthmap.assign(width * height, 0.0f);
    float z = 0.0f;
    for(size_t i = 0; i < width * height; ++i)
    {
      f >> z;
      depthmap[i] = z / 100.0f;
    }
    return true;
  }

  std::vector<DatasetEntry>::const_iterator begin() const { return dataset_.begin(); }
  std::vector<DatasetEntry>::const_iterator end() const { return dataset_.end(); }
  const DatasetEntry &operator()(size_t index) const { return dataset_.at(index); }
  size_t size() const { return dataset_.size(); }

private:
  static std::string join(const std::string &a, const std::string &b)
  {
    return (a.empty() || a[a.size() - 1] == '/') ? a + b : a + "/" + b;
  }
  std::string dataset_path_, sequence_file_;
  std::vector<DatasetEntry> dataset_;
};

} // namespace test
} // namespace rmd

#endif // RMD_DATASET_READER_H
