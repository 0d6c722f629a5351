// stub: gipuma.cu:28-30 includes this header but uses no cv:: symbol (test-infrastructure only)
namespace cv {}
