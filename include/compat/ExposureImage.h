// ExposureImage — drop-in for the reference's value type (src/ExposureImage.h:33-51):
// an owning float image plus frame metadata, returned by DatasetReader::getImage and
// deleted by the caller.  Field names, constructor signature and new[]/delete[] ownership
// are part of the surface downstream code (DSO, the calibration tools) relies on.
#pragma once

class ExposureImage
{
public:
	float* image;          // w*h floats, row-major, owned
	double timestamp;
	int w, h;
	float exposure_time;   // milliseconds (times.txt column 3), 0 if unknown
	int id;

	ExposureImage(int width, int height, double stamp, float exposure, int frameId)
		: image(new float[(long)width * height]), timestamp(stamp), w(width), h(height),
		  exposure_time(exposure), id(frameId) {}
	~ExposureImage() { delete[] image; }

private:
	ExposureImage(const ExposureImage&);             // owning raw pointer: not copyable
	ExposureImage& operator=(const ExposureImage&);
};
