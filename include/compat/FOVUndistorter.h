// UndistorterFOV — drop-in for the reference's FOV-model rectifier (src/FOVUndistorter.h:36-96,
// src/FOVUndistorter.cpp) on top of the B200 C ABI (include/mdc_b200.h).  Header-only: it is
// compiled in the consumer's translation unit against the consumer's own Eigen, and talks to
// libmdc_b200.so through plain C calls only.
//
// Same public surface and behaviour:
//   * constructor parses camera.txt, prints the same diagnostics, leaves isValid()==false on any
//     format problem or "none" (FOVUndistorter.cpp:55-126);
//   * remap tables are built on the host, bit-identical to the reference (FOVUndistorter.cpp:224-251);
//   * undistort<T> (T = float, unsigned char; FOVUndistorter.cpp:322-370) takes HOST pointers,
//     returns without touching the output for an invalid object or wrong pixel counts — but the
//     bilinear remap itself runs in the sm_100a kernel (H2D copy, kernel, D2H copy; no CPU path).
// The device is chosen with the MDC_DEVICE environment variable (default 0).
#pragma once
#include <cstdio>
#include <cstdlib>

#include "Eigen/Core"
#include "ExposureImage.h"
#include "mdc_b200.h"

class UndistorterFOV
{
public:
	UndistorterFOV(const char* configFileName) : model(0), context(0) { mdc_fov_create(configFileName, &model); }
	UndistorterFOV() : model(0), context(0) {}
	~UndistorterFOV()
	{
		if(context != 0) mdc_ctx_destroy(context);
		if(model != 0) mdc_fov_destroy(model);
	}

	template<typename T>
	void undistort(const T* input, float* output, int nPixIn, int nPixOut) const;
	void distortCoordinates(float* in_x, float* in_y, int n)
	{
		if(model == 0) { printf("ERROR: invalid UndistorterFOV!\n"); return; }
		mdc_fov_distort_coordinates(model, in_x, in_y, n);   // prints the same error itself when invalid
	}

	inline Eigen::Matrix3f getK_rect() const { return matrixOf(true); }
	inline Eigen::Matrix3f getK_org() const { return matrixOf(false); }
	inline float getOmega() const { return mdc_fov_omega(model); }
	const Eigen::VectorXf getOriginalCalibration() const
	{
		float v[5] = {0,0,0,0,0};
		if(model != 0) mdc_fov_original_calibration(model, v);
		Eigen::VectorXf vec(5);
		for(int i=0;i<5;i++) vec[i] = v[i];
		return vec;
	}
	const Eigen::Vector2i getInputDims() const
	{
		int iw=0, ih=0;
		if(model != 0) mdc_fov_dims(model, &iw, &ih, 0, 0);
		return Eigen::Vector2i(iw, ih);
	}
	const Eigen::Vector2i getOutputDims() const
	{
		int ow=0, oh=0;
		if(model != 0) mdc_fov_dims(model, 0, 0, &ow, &oh);
		return Eigen::Vector2i(ow, oh);
	}
	bool isValid() const { return model != 0 && mdc_fov_is_valid(model) != 0; }

	// --- additions (not in the reference): access for DatasetReader's fused device path
	const mdc_fov* b200Model() const { return model; }
	static int b200Device() { const char* e = getenv("MDC_DEVICE"); return e ? atoi(e) : 0; }

private:
	UndistorterFOV(const UndistorterFOV&);
	UndistorterFOV& operator=(const UndistorterFOV&);

	Eigen::Matrix3f matrixOf(bool rect) const
	{
		float kr[9] = {0,0,0,0,0,0,0,0,0}, ko[9] = {0,0,0,0,0,0,0,0,0};
		if(model != 0) mdc_fov_get_K(model, kr, ko);
		const float* k = rect ? kr : ko;
		Eigen::Matrix3f K;
		for(int r=0;r<3;r++) for(int c=0;c<3;c++) K(r,c) = k[3*r+c];
		return K;
	}
	mdc_ctx* deviceContext() const
	{
		if(context == 0 && mdc_ctx_create(b200Device(), model, 0, &context) != MDC_OK)
		{
			printf("UndistorterFOV: cannot create the B200 device context: %s\n", mdc_last_error());
			context = 0;
		}
		return context;
	}

	mdc_fov* model;
	mutable mdc_ctx* context;   // created on first undistort(); tables stay resident in HBM
};

// The reference instantiates exactly these two (FOVUndistorter.cpp:369-370).
template<> inline void UndistorterFOV::undistort<float>(const float* input, float* output, int nPixIn, int nPixOut) const
{
	if(!isValid()) return;
	mdc_ctx* c = deviceContext();
	if(c != 0 && mdc_undistort_f32_host(c, input, output, nPixIn, nPixOut) == MDC_ERR_CUDA)
		printf("UndistorterFOV::undistort<float>: %s\n", mdc_last_error());
}
template<> inline void UndistorterFOV::undistort<unsigned char>(const unsigned char* input, float* output, int nPixIn, int nPixOut) const
{
	if(!isValid()) return;
	mdc_ctx* c = deviceContext();
	if(c != 0 && mdc_undistort_u8_host(c, input, output, nPixIn, nPixOut) == MDC_ERR_CUDA)
		printf("UndistorterFOV::undistort<unsigned char>: %s\n", mdc_last_error());
}
