// DatasetReader — drop-in for the reference's sequence reader (src/BenchmarkDatasetReader.h:44-345).
//
// File handling (images/ folder or images.zip via libzip, times.txt, cv::imread / cv::imdecode) is
// ordinary host code and stays as the reference has it; image decode is out of scope for the B200
// path.  What changes is getImage(): instead of running unMapImage into a
// temporary float image and then undistort<float> on the CPU (:222-223), the raw 8-bit frame goes to
// the GPU once and the whole mode switch of :210-241 is one fused sm_100a kernel
// (mdc_prepare_batch_host), bit-identical to the reference's result.
//
// Same public surface: getdir(), DatasetReader(folder), getUndistorter(), getPhotoUndistorter(),
// getNumImages(), getTimestamp(), getExposure(), getImage(), getImageRaw_internal().
#pragma once
#include <sstream>
#include <fstream>
#include <dirent.h>
#include <algorithm>
#include <cassert>
#include <string>
#include <vector>

#include "opencv2/opencv.hpp"
#include "FOVUndistorter.h"
#include "PhotometricUndistorter.h"

#include "zip.h"

// Lists `dir` (sorted, full paths) into `files`; -1 if the directory cannot be opened.
inline int getdir(std::string dir, std::vector<std::string> &files)
{
	DIR* handle = opendir(dir.c_str());
	if(handle == NULL) return -1;
	for(struct dirent* entry = readdir(handle); entry != NULL; entry = readdir(handle))
	{
		const std::string name(entry->d_name);
		if(name == "." || name == "..") continue;
		files.push_back(name);
	}
	closedir(handle);
	std::sort(files.begin(), files.end());

	if(dir.empty() || dir[dir.length()-1] != '/') dir += "/";
	for(size_t i=0;i<files.size();i++)
		if(files[i].at(0) != '/') files[i] = dir + files[i];
	return (int)files.size();
}

class DatasetReader
{
public:
	DatasetReader(std::string folder)
		: path(folder), isZipped(false), undistorter(0), photoUndistorter(0), ziparchive(0), databuffer(0), deviceContext(0)
	{
		getdir(path+"images/", files);
		if(!files.empty())
		{
			printf("Load Dataset %s: found %d files in folder /images; assuming that all images are there.\n",
					path.c_str(), (int)files.size());
		}
		else
		{
			printf("Load Dataset %s: found no in folder /images; assuming that images are zipped.\n", path.c_str());
			isZipped = true;
			int ziperror = 0;
			ziparchive = zip_open((path+"images.zip").c_str(), ZIP_RDONLY, &ziperror);
			if(ziperror != 0)
			{
				printf("ERROR %d reading archive %s!\n", ziperror, (path+"images.zip").c_str());
				exit(1);
			}
			files.clear();
			const int numEntries = (int)zip_get_num_entries(ziparchive, 0);
			for(int k=0;k<numEntries;k++)
			{
				const std::string entry(zip_get_name(ziparchive, k, ZIP_FL_ENC_STRICT));
				if(entry == "." || entry == "..") continue;
				files.push_back(entry);
			}
			printf("got %d entries and %d files from zipfile!\n", numEntries, (int)files.size());
			std::sort(files.begin(), files.end());
		}
		loadTimestamps(path+"times.txt");

		// calibration models (host) ...
		undistorter = new UndistorterFOV((path+"camera.txt").c_str());
		photoUndistorter = new PhotometricUndistorter(path+"pcalib.txt", path+"vignette.png",
				undistorter->getInputDims()[0], undistorter->getInputDims()[1]);
		widthOrg = undistorter->getInputDims()[0];
		heightOrg = undistorter->getInputDims()[1];
		width = undistorter->getOutputDims()[0];
		height = undistorter->getOutputDims()[1];

		// ... and one device context holding all four tables for the fused getImage kernel
		if(mdc_ctx_create(UndistorterFOV::b200Device(), undistorter->b200Model(), photoUndistorter->b200Model(), &deviceContext) != MDC_OK)
		{
			printf("DatasetReader: cannot create the B200 device context: %s\n", mdc_last_error());
			deviceContext = 0;
		}
		printf("Dataset %s: Got %d files!\n", path.c_str(), (int)getNumImages());
	}
	~DatasetReader()
	{
		if(deviceContext != 0) mdc_ctx_destroy(deviceContext);
		delete undistorter;
		delete photoUndistorter;
		if(ziparchive != 0) zip_close(ziparchive);
		delete[] databuffer;
	}

	UndistorterFOV* getUndistorter() { return undistorter; }
	PhotometricUndistorter* getPhotoUndistorter() { return photoUndistorter; }
	int getNumImages() { return (int)files.size(); }
	double getTimestamp(int id) { return (id < 0 || id >= (int)timestamps.size()) ? 0 : timestamps[id]; }
	float getExposure(int id) { return (id < 0 || id >= (int)exposures.size()) ? 0 : exposures[id]; }

	// Returns a heap ExposureImage the caller deletes, or 0 if the decoded frame has the wrong size/type.
	ExposureImage* getImage(int id, bool rectify, bool removeGamma, bool removeVignette, bool nanOverexposed)
	{
		assert(id >= 0 && id < (int)files.size());
		cv::Mat imageRaw = getImageRaw_internal(id);
		if(imageRaw.rows != heightOrg || imageRaw.cols != widthOrg)
		{
			printf("ERROR: expected cv-mat to have dimensions %d x %d; found %d x %d (image %s)!\n",
					widthOrg, heightOrg, imageRaw.cols, imageRaw.rows, files[id].c_str());
			return 0;
		}
		if(imageRaw.type() != CV_8U)
		{
			printf("ERROR: expected cv-mat to have type 8U!\n");
			return 0;
		}

		const bool rectified = rectify;
		ExposureImage* ret = new ExposureImage(rectified ? width : widthOrg, rectified ? height : heightOrg,
				timestamps[id], exposures[id], id);
		const unsigned flags = (rectify ? MDC_RECTIFY : 0u) | (removeGamma ? MDC_REMOVE_GAMMA : 0u)
				| (removeVignette ? MDC_REMOVE_VIGNETTE : 0u) | (nanOverexposed ? MDC_NAN_OVEREXPOSED : 0u);
		float* levels[1] = { ret->image };
		int status = deviceContext != 0 ? mdc_prepare_batch_host(deviceContext, imageRaw.data, 1, flags, levels, 1) : MDC_ERR_CUDA;
		if(status == MDC_ERR_INVALID_OBJECT)
			return ret;   // invalid rectifier: like the reference, the image is returned unwritten (undistort is a no-op)
		if(status != MDC_OK)
			printf("DatasetReader::getImage: %s\n", deviceContext != 0 ? mdc_last_error() : "no device context");
		return ret;
	}

	cv::Mat getImageRaw_internal(int id)
	{
		if(!isZipped)
			return cv::imread(files[id], CV_LOAD_IMAGE_GRAYSCALE);

		// zipped: inflate into a scratch buffer (grown once if the first guess is too small), then decode
		long capacity = (long)widthOrg*heightOrg*6+10000;
		if(databuffer == 0) databuffer = new char[capacity];
		zip_file_t* entry = zip_fopen(ziparchive, files[id].c_str(), 0);
		long readbytes = zip_fread(entry, databuffer, capacity);
		if(readbytes > (long)widthOrg*heightOrg*6)
		{
			printf("read %ld/%ld bytes for file %s. increase buffer!!\n", readbytes, capacity, files[id].c_str());
			delete[] databuffer;
			capacity = (long)widthOrg*heightOrg*60+1000000;
			databuffer = new char[capacity];
			entry = zip_fopen(ziparchive, files[id].c_str(), 0);
			readbytes = zip_fread(entry, databuffer, capacity-900000);
			if(readbytes > capacity-990000)
			{
				printf("buffer still to small (read %ld/%ld). abort.\n", readbytes, capacity-900000);
				exit(1);
			}
		}
		return cv::imdecode(cv::Mat((int)readbytes, 1, CV_8U, databuffer), CV_LOAD_IMAGE_GRAYSCALE);
	}

private:
	// times.txt: "id stamp [exposure_ms]" per line; on a count mismatch everything is zeroed.
	inline void loadTimestamps(std::string timesFile)
	{
		timestamps.clear();
		exposures.clear();
		std::ifstream in(timesFile.c_str());
		std::string line;
		while(std::getline(in, line))
		{
			int id; double stamp; float exposure = 0;
			const int got = sscanf(line.c_str(), "%d %lf %f", &id, &stamp, &exposure);
			if(got < 2) continue;
			timestamps.push_back(stamp);
			exposures.push_back(got == 3 ? exposure : 0);
		}
		if((int)exposures.size() != getNumImages())
		{
			printf("DatasetReader: Mismatch between number of images and number of timestamps / exposure times. Set all to zero.");
			timestamps.assign(getNumImages(), 0.0);
			exposures.assign(getNumImages(), 0.0f);
		}
	}

	std::vector<std::string> files;
	std::vector<double> timestamps;
	std::vector<float> exposures;
	int width, height;
	int widthOrg, heightOrg;
	std::string path;
	bool isZipped;

	UndistorterFOV* undistorter;
	PhotometricUndistorter* photoUndistorter;
	zip_t* ziparchive;
	char* databuffer;
	mdc_ctx* deviceContext;
};
