// Implementation of the stand-in cv:: free functions (test infrastructure only).
#include "opencv2/core/core.hpp"
#include <cmath>
#include <map>
#include <fstream>

namespace {
struct Registered { int rows, cols, type; std::vector<unsigned char> px; };
std::map<std::string, Registered>& registry() { static std::map<std::string, Registered> r; return r; }

// binary PGM ("P5"), maxval < 256 -> CV_8U, else CV_16U (big-endian on disk)
bool read_pgm(const std::string& path, cv::Mat& out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.good()) return false;
    std::string magic; f >> magic;
    if (magic != "P5") return false;
    int vals[3], got = 0;
    while (got < 3 && f.good()) {
        int c = f.peek();
        if (c == '#') { std::string skip; std::getline(f, skip); continue; }
        if (isspace(c)) { f.get(); continue; }
        f >> vals[got++];
    }
    if (got != 3) return false;
    f.get();  // single whitespace after maxval
    int w = vals[0], h = vals[1], maxv = vals[2];
    if (maxv < 256) {
        out = cv::Mat(h, w, CV_8U);
        f.read((char*)out.data, (std::streamsize)w * h);
    } else {
        out = cv::Mat(h, w, CV_16U);
        std::vector<unsigned char> raw((size_t)w * h * 2);
        f.read((char*)&raw[0], (std::streamsize)raw.size());
        for (size_t i = 0; i < (size_t)w * h; i++)
            out.at<ushort>((int)i) = (ushort)((raw[2 * i] << 8) | raw[2 * i + 1]);
    }
    return true;
}
}  // namespace

extern "C" void mdc_shim_register_image(const char* path, int rows, int cols, int type, const void* pixels) {
    Registered r; r.rows = rows; r.cols = cols; r.type = type;
    size_t bytes = (size_t)rows * cols * cv::shim_elem_size(type);
    r.px.assign((const unsigned char*)pixels, (const unsigned char*)pixels + bytes);
    registry()[path] = r;
}
extern "C" void mdc_shim_clear_images() { registry().clear(); }

namespace cv {
Mat imread(const std::string& path, int flags) {
    std::map<std::string, Registered>::iterator it = registry().find(path);
    if (it != registry().end()) {
        Mat m(it->second.rows, it->second.cols, it->second.type);
        if (!it->second.px.empty()) memcpy(m.data, &it->second.px[0], it->second.px.size());
        return m;
    }
    Mat m;
    if (read_pgm(path, m)) {
        if (flags == CV_LOAD_IMAGE_GRAYSCALE && m.type() == CV_16U) {  // grayscale load narrows to 8 bit
            Mat n(m.rows, m.cols, CV_8U);
            for (int i = 0; i < m.rows * m.cols; i++) n.at<uchar>(i) = (uchar)(m.at<ushort>(i) >> 8);
            return n;
        }
        return m;
    }
    return Mat();
}
Mat imdecode(const Mat&, int) { return Mat(); }
// No encoder here.  With MDC_SHIM_DUMP_DIR set, the pixels are dumped raw instead (int32 rows, cols, type; then the data) under the
// file's base name, so that a test can look at what a reference program wanted to save.
bool imwrite(const std::string& path, const Mat& m) {
    const char* dir = getenv("MDC_SHIM_DUMP_DIR");
    if (!dir || !m.data) return true;
    const size_t slash = path.find_last_of('/');
    const std::string out = std::string(dir) + "/" + (slash == std::string::npos ? path : path.substr(slash + 1)) + ".raw";
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) return false;
    const int head[3] = {m.rows, m.cols, m.type()};
    fwrite(head, sizeof head, 1, f);
    fwrite(m.data, (size_t)m.rows * m.cols * shim_elem_size(m.type()), 1, f);
    fclose(f);
    return true;
}
void imshow(const std::string&, const Mat&) {}
void Mat::convertTo(Mat& dst, int type, double alpha, double beta) const {
    dst = Mat(rows, cols, type);
    const size_t n = (size_t)rows * cols;
    for (size_t i = 0; i < n; i++) {
        const double v = nearbyint((double)((const float*)data)[i] * alpha + beta);      // NaN / negative -> 0, like saturate_cast
        if (type == CV_8U) dst.at<uchar>((int)i) = (uchar)(v > 255 ? 255 : (v > 0 ? v : 0));
        else if (type == CV_16U) dst.at<ushort>((int)i) = (ushort)(v > 65535 ? 65535 : (v > 0 ? v : 0));
    }
}
Mat findHomography(const std::vector<Point2f>&, const std::vector<Point2f>&) {
    static std::ifstream f;
    static bool opened = false;
    if (!opened) {
        opened = true;
        const char* path = getenv("MDC_SHIM_HOMOGRAPHIES");
        if (path) f.open(path);
    }
    Mat h(3, 3, CV_64F);
    for (int i = 0; i < 9; i++) {
        double v = (i % 4 == 0) ? 1.0 : 0.0;
        if (f.good()) f >> v;
        h.at<double>(i) = v;
    }
    return h;
}
int waitKey(int) { return ' '; }
}  // namespace cv
