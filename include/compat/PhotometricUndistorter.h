// PhotometricUndistorter — drop-in for the reference's photometric un-mapper
// (src/PhotometricUndistorter.h:37-54, src/PhotometricUndistorter.cpp) on the B200 C ABI.
// Header-only; see FOVUndistorter.h for the conventions.
//
//   * constructor: parses pcalib.txt (256 strictly increasing floats, normalised to 0..255) and the
//     vignette image (8/16-bit grey PNG or PGM) with the reference's messages and validity rules
//     (PhotometricUndistorter.cpp:56-156); tables are bit-identical to the reference's;
//   * unMapImage (PhotometricUndistorter.cpp:165-212): HOST pointers in/out, same flag
//     sanitising, computed by the sm_100a kernel (no CPU path);
//   * getGInv()/getG() return 0 while the response is not loaded (PhotometricUndistorter.h:44-45).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <string>

#include "mdc_b200.h"

class PhotometricUndistorter
{
public:
	PhotometricUndistorter(std::string file, std::string vignetteImage, int w_, int h_) : model(0), context(0)
	{
		mdc_photo_create(file.c_str(), vignetteImage.c_str(), w_, h_, &model);
	}
	~PhotometricUndistorter()
	{
		if(context != 0) mdc_ctx_destroy(context);
		if(model != 0) mdc_photo_destroy(model);
	}

	void unMapImage(unsigned char* image_in, float* image_out, int n, bool undoGamma, bool undoVignette, bool killOverexposed)
	{
		if(context == 0)
		{
			const char* e = getenv("MDC_DEVICE");
			if(mdc_ctx_create(e ? atoi(e) : 0, 0, model, &context) != MDC_OK)
			{
				printf("PhotometricUndistorter: cannot create the B200 device context: %s\n", mdc_last_error());
				context = 0;
				return;
			}
		}
		unsigned flags = (undoGamma ? MDC_REMOVE_GAMMA : 0u) | (undoVignette ? MDC_REMOVE_VIGNETTE : 0u)
				| (killOverexposed ? MDC_NAN_OVEREXPOSED : 0u);
		if(mdc_unmap_u8_host(context, image_in, image_out, n, flags) != MDC_OK)
			printf("PhotometricUndistorter::unMapImage: %s\n", mdc_last_error());
	}
	float* getGInv() { return mdc_photo_ginv(model); }
	float* getG() { return mdc_photo_g(model); }

	// --- addition (not in the reference): access for DatasetReader's fused device path
	const mdc_photo* b200Model() const { return model; }

private:
	PhotometricUndistorter(const PhotometricUndistorter&);
	PhotometricUndistorter& operator=(const PhotometricUndistorter&);

	mdc_photo* model;
	mdc_ctx* context;
};
