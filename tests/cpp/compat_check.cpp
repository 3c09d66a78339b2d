// Drives the C++ drop-in classes (include/compat/*.h) the way the reference's programs do and dumps
// every result to a binary file that tests/test_compat_cpp.py compares with the CPU oracle.
//   compat_check <dataset_dir/> <out.bin>
// Layout of out.bin: for each image id, for flags 0..15: int32 w, int32 h, float[w*h] (getImage);
// then for image 0: undistort<uchar>, undistort<float>(unMapImage(1,1,1)), and the getters.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "BenchmarkDatasetReader.h"

static void put(FILE* f, const void* p, size_t n) { if (fwrite(p, 1, n, f) != n) { perror("fwrite"); exit(2); } }

int main(int argc, char** argv)
{
	if(argc < 3) { printf("usage: compat_check <dataset_dir/> <out.bin>\n"); return 1; }
	DatasetReader* reader = new DatasetReader(argv[1]);
	FILE* f = fopen(argv[2], "wb");
	if(!f) { perror("fopen"); return 2; }
	const int n = reader->getNumImages();
	put(f, &n, 4);
	for(int id=0; id<n; id++)
		for(int flags=0; flags<16; flags++)
		{
			ExposureImage* img = reader->getImage(id, flags&1, (flags>>1)&1, (flags>>2)&1, (flags>>3)&1);
			if(img == 0) { int z[2] = {0,0}; put(f, z, 8); continue; }
			put(f, &img->w, 4); put(f, &img->h, 4);
			put(f, img->image, sizeof(float)*img->w*img->h);
			if(img->id != id) return 3;
			delete img;
		}
	// stand-alone operators on frame 0
	UndistorterFOV* u = reader->getUndistorter();
	PhotometricUndistorter* p = reader->getPhotoUndistorter();
	const int iw = u->getInputDims()[0], ih = u->getInputDims()[1], ow = u->getOutputDims()[0], oh = u->getOutputDims()[1];
	cv::Mat raw = reader->getImageRaw_internal(0);
	std::vector<float> tmp(iw*ih), out(ow*oh, -1.0f);
	u->undistort<unsigned char>(raw.data, &out[0], iw*ih, ow*oh);
	put(f, &out[0], sizeof(float)*out.size());
	p->unMapImage(raw.data, &tmp[0], iw*ih, true, true, true);
	u->undistort<float>(&tmp[0], &out[0], iw*ih, ow*oh);
	put(f, &out[0], sizeof(float)*out.size());
	// wrong pixel count: output must stay untouched
	std::vector<float> keep(ow*oh, 5.0f);
	u->undistort<float>(&tmp[0], &keep[0], iw*ih-1, ow*oh);
	put(f, &keep[0], sizeof(float)*keep.size());
	// getters
	Eigen::Matrix3f K = u->getK_rect(), Ko = u->getK_org();
	float g[18];
	for(int r=0;r<3;r++) for(int c=0;c<3;c++) { g[3*r+c] = K(r,c); g[9+3*r+c] = Ko(r,c); }
	put(f, g, sizeof g);
	float om = u->getOmega(); put(f, &om, 4);
	Eigen::VectorXf oc = u->getOriginalCalibration();
	for(int i=0;i<5;i++) { float v = oc[i]; put(f, &v, 4); }
	put(f, p->getGInv(), 256*4);
	double ts = reader->getTimestamp(1); float ex = reader->getExposure(1);
	put(f, &ts, 8); put(f, &ex, 4);
	float xs[3] = {0.0f, 10.5f, (float)ow}, ys[3] = {0.0f, 20.25f, (float)oh};
	u->distortCoordinates(xs, ys, 3);
	put(f, xs, 12); put(f, ys, 12);
	fclose(f);
	delete reader;
	printf("compat_check done\n");
	return 0;
}
