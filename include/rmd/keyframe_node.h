// rmd/keyframe_node.h -- the keyframe state machine of rmd::DepthmapNode
// (include/rmd/depthmap_node.h:30-62, src/depthmap_node.cpp:88-183) without ROS:
// the caller of the depth-filter path, as host-side C++
// above the drop-in headers.  `DepthmapT` is rmd::Depthmap (or anything with its
// interface: setReferenceImage, update, getConvergedPercentage, getDistFromRef,
// downloadDenoisedDepthmap, downloadConvergenceMap); `ImageT` is what it takes as
// a frame (cv::Mat in the reference); `PublisherT` receives the two publish
// requests of the node (rmd::Publisher::publishDepthmapAndPointCloud /
// publishConvergenceMap).  No OpenCV / ROS headers are needed to compile this.
#ifndef RMD_KEYFRAME_NODE_H
#define RMD_KEYFRAME_NODE_H

#include <rmd/se3.cuh>

namespace rmd
{

namespace ProcessingStates
{
enum State   // include/rmd/depthmap_node.h:32-36
{
  UPDATE,
  TAKE_REFERENCE_FRAME
};
}
typedef ProcessingStates::State State;

template<typename DepthmapT, typename ImageT, typename PublisherT>
class KeyframeNode
{
public:
  // Parameter defaults of DepthmapNode::init, src/depthmap_node.cpp:81-83.
  KeyframeNode(DepthmapT &depthmap, PublisherT &publisher, float ref_compl_perc = 10.0f,
               float max_dist_from_ref = 0.5f, int publish_conv_every_n = 10)
    : depthmap_(depthmap)
    , publisher_(publisher)
    , state_(State::TAKE_REFERENCE_FRAME)        // src/depthmap_node.cpp:35
    , ref_compl_perc_(ref_compl_perc)
    , max_dist_from_ref_(max_dist_from_ref)
    , publish_conv_every_n_(publish_conv_every_n)
    , num_msgs_(0)
  {
  }

  // DepthmapNode::denseInputCallback, src/depthmap_node.cpp:88-162: one
  // svo_msgs/DenseInput = 8-bit image, camera pose in the world, scene depth range.
  void denseInput(const ImageT &img_8uc1, const SE3<float> &T_world_curr, float min_depth, float max_depth)
  {
    num_msgs_ += 1;
    switch(state_)
    {
    case State::TAKE_REFERENCE_FRAME:
      if(depthmap_.setReferenceImage(img_8uc1, T_world_curr.inv(), min_depth, max_depth))
        state_ = State::UPDATE;
      break;
    case State::UPDATE:
    {
      depthmap_.update(img_8uc1, T_world_curr.inv());
      const float perc_conv = depthmap_.getConvergedPercentage();
      const float dist_from_ref = depthmap_.getDistFromRef();
      if(perc_conv > ref_compl_perc_ || dist_from_ref > max_dist_from_ref_)
      {
        state_ = State::TAKE_REFERENCE_FRAME;
        denoiseAndPublishResults();
      }
      break;
    }
    default:
      break;
    }
    if(publish_conv_every_n_ < num_msgs_)
    {
      publishConvergenceMap();
      num_msgs_ = 0;
    }
  }

  State state() const { return state_; }
  int numMsgs() const { return num_msgs_; }

private:
  void denoiseAndPublishResults()               // src/depthmap_node.cpp:165-173
  {
    depthmap_.downloadDenoisedDepthmap(0.5f, 200);
    depthmap_.downloadConvergenceMap();
    publisher_.publishDepthmapAndPointCloud();
  }

  void publishConvergenceMap()                  // src/depthmap_node.cpp:175-183
  {
    depthmap_.downloadConvergenceMap();
    publisher_.publishConvergenceMap();
  }

  DepthmapT &depthmap_;
  PublisherT &publisher_;
  State state_;
  float ref_compl_perc_;
  float max_dist_from_ref_;
  int publish_conv_every_n_;
  int num_msgs_;
};

} // namespace rmd

#endif // RMD_KEYFRAME_NODE_H
