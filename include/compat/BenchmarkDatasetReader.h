// DatasetReader — drop-in for the reference's sequence reader (src/BenchmarkDatasetReader.h:44-345).
//
// Public surface kept: getdir(), DatasetReader(folder), getUndistorter(), getPhotoUndistorter(), getNumImages(),
// getTimestamp(), getExposure(), getImage(), getImageRaw_internal().  File handling (an images/ folder or
// images.zip through libzip, times.txt, cv::imread / cv::imdecode) is ordinary host code; image decode is out
// of scope for the B200 path.
//
// What changes is getImage(): the reference runs unMapImage into a temporary float image and then
// undistort<float> on the CPU (:222-223).  Here the raw 8-bit frame goes to the GPU once and the whole mode
// switch of :210-241 is ONE fused sm_100a kernel launch (mdc_prepare_batch_host) whose result is bit-identical
// to the reference's; all four calibration tables live in one device context owned by the reader.
//
// Define MDC_NATIVE_SEQUENCE_READER before including this header to drop the libzip and cv::imread/imdecode dependencies as
// well: folder listing, images.zip, times.txt and frame decode (baseline JPEG, grey PNG, PGM -> the bytes cv::imread would
// return) then come from libmdc_b200's own sequence reader (mdc_seq_*); cv::Mat is still the frame type.
#pragma once
#include <algorithm>
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <dirent.h>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "opencv2/opencv.hpp"
#include "FOVUndistorter.h"
#include "PhotometricUndistorter.h"

#ifndef MDC_NATIVE_SEQUENCE_READER
#include "zip.h"
#endif

// Sorted list of the entries of `dir` (full paths appended to `files`); -1 if it cannot be opened.
inline int getdir(std::string dir, std::vector<std::string> &files)
{
	DIR* handle = opendir(dir.c_str());
	if(handle == NULL) return -1;
	while(struct dirent* entry = readdir(handle))
	{
		const std::string name(entry->d_name);
		if(name != "." && name != "..") files.push_back(name);
	}
	closedir(handle);
	std::sort(files.begin(), files.end());

	if(dir.empty() || dir[dir.size()-1] != '/') dir.push_back('/');
	for(std::vector<std::string>::iterator it = files.begin(); it != files.end(); ++it)
		if((*it)[0] != '/') *it = dir + *it;
	return (int)files.size();
}

class DatasetReader
{
public:
	DatasetReader(std::string folder)
		: width(0), height(0), widthOrg(0), heightOrg(0), path(folder), isZipped(false),
		  undistorter(0), photoUndistorter(0), ziparchive(0), deviceContext(0)
	{
#ifdef MDC_NATIVE_SEQUENCE_READER
		if(mdc_seq_open(path.c_str(), &ziparchive) != MDC_OK) exit(1);      // the reference exits when images.zip cannot be read
		isZipped = mdc_seq_is_zipped(ziparchive) != 0;
		for(int i = 0; i < mdc_seq_num_images(ziparchive); i++)
		{
			files.push_back(mdc_seq_name(ziparchive, i));
			timestamps.push_back(mdc_seq_timestamp(ziparchive, i));
			exposures.push_back(mdc_seq_exposure(ziparchive, i));
		}
		loadCalibration();
#else
		locateImages();
		loadTimestamps(path + "times.txt");
		loadCalibration();
		printf("Dataset %s: Got %d files!\n", path.c_str(), getNumImages());
#endif
	}
	~DatasetReader()
	{
		if(deviceContext != 0) mdc_ctx_destroy(deviceContext);
		delete undistorter;
		delete photoUndistorter;
#ifdef MDC_NATIVE_SEQUENCE_READER
		mdc_seq_close(ziparchive);
#else
		if(ziparchive != 0) zip_close(ziparchive);
#endif
	}

	UndistorterFOV* getUndistorter() { return undistorter; }
	PhotometricUndistorter* getPhotoUndistorter() { return photoUndistorter; }
	int getNumImages() { return (int)files.size(); }
	double getTimestamp(int id) { return inRange(id, timestamps.size()) ? timestamps[id] : 0; }
	float getExposure(int id) { return inRange(id, exposures.size()) ? exposures[id] : 0; }

	// Heap ExposureImage the caller deletes; 0 if the decoded frame has the wrong size or type.
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

		ExposureImage* result = new ExposureImage(rectify ? width : widthOrg, rectify ? height : heightOrg,
				timestamps[id], exposures[id], id);
		unsigned mode = 0;
		if(rectify) mode |= MDC_RECTIFY;
		if(removeGamma) mode |= MDC_REMOVE_GAMMA;
		if(removeVignette) mode |= MDC_REMOVE_VIGNETTE;
		if(nanOverexposed) mode |= MDC_NAN_OVEREXPOSED;
		float* level0[1] = { result->image };
		const int status = (deviceContext != 0)
				? mdc_prepare_batch_host(deviceContext, imageRaw.data, 1, mode, level0, 1) : MDC_ERR_CUDA;
		// an invalid rectifier leaves the image unwritten, like the reference (undistort is then a no-op)
		if(status != MDC_OK && status != MDC_ERR_INVALID_OBJECT)
		{
			// no device / CUDA failure: the reference has no such state; never hand out an image that was not computed
			printf("DatasetReader::getImage: %s\n", deviceContext != 0 ? mdc_last_error() : "no device context");
			delete result;
			return 0;
		}
		return result;
	}

	cv::Mat getImageRaw_internal(int id)
	{
#ifdef MDC_NATIVE_SEQUENCE_READER
		// frames normally have the calibrated size: decode straight into a Mat of that size, re-decode only if it differs
		int w = 0, h = 0;
		cv::Mat frame(heightOrg, widthOrg, CV_8U);
		int status = mdc_seq_read_gray8(ziparchive, id, frame.data, (size_t)widthOrg * heightOrg, &w, &h);
		if(status == MDC_OK && w == widthOrg && h == heightOrg) return frame;
		if(w < 1 || h < 1) return cv::Mat();                 // undecodable: empty, like cv::imread
		cv::Mat other(h, w, CV_8U);
		status = mdc_seq_read_gray8(ziparchive, id, other.data, (size_t)w * h, &w, &h);
		return status == MDC_OK ? other : cv::Mat();
#else
		if(!isZipped) return cv::imread(files[id], CV_LOAD_IMAGE_GRAYSCALE);
		long bytes = readArchiveEntry(id);
		return cv::imdecode(cv::Mat((int)bytes, 1, CV_8U, &databuffer[0]), CV_LOAD_IMAGE_GRAYSCALE);
#endif
	}

private:
	static bool inRange(int id, size_t n) { return id >= 0 && id < (int)n; }
#ifndef MDC_NATIVE_SEQUENCE_READER

	// images/ folder if it has entries, else images.zip (exit(1) if that cannot be opened, like the reference)
	void locateImages()
	{
		getdir(path + "images/", files);
		if(!files.empty())
		{
			printf("Load Dataset %s: found %d files in folder /images; assuming that all images are there.\n",
					path.c_str(), (int)files.size());
			return;
		}
		printf("Load Dataset %s: found no in folder /images; assuming that images are zipped.\n", path.c_str());
		isZipped = true;
		const std::string archive = path + "images.zip";
		int ziperror = 0;
		ziparchive = zip_open(archive.c_str(), ZIP_RDONLY, &ziperror);
		if(ziperror != 0)
		{
			printf("ERROR %d reading archive %s!\n", ziperror, archive.c_str());
			exit(1);
		}
		const int numEntries = (int)zip_get_num_entries(ziparchive, 0);
		for(int k = 0; k < numEntries; k++)
		{
			const std::string entry(zip_get_name(ziparchive, k, ZIP_FL_ENC_STRICT));
			if(entry != "." && entry != "..") files.push_back(entry);
		}
		printf("got %d entries and %d files from zipfile!\n", numEntries, (int)files.size());
		std::sort(files.begin(), files.end());
	}

	// inflate entry `id` into databuffer; first guess 6 bytes/pixel, one retry at 60 bytes/pixel, then give up
	long readArchiveEntry(int id)
	{
		const long pixels = (long)widthOrg * heightOrg;
		if(databuffer.empty()) databuffer.resize(pixels*6 + 10000);
		zip_file_t* entry = zip_fopen(ziparchive, files[id].c_str(), 0);
		long got = zip_fread(entry, &databuffer[0], (long)databuffer.size());
		if(got > pixels*6)
		{
			printf("read %ld/%ld bytes for file %s. increase buffer!!\n", got, (long)databuffer.size(), files[id].c_str());
			databuffer.assign(pixels*60 + 1000000, 0);
			entry = zip_fopen(ziparchive, files[id].c_str(), 0);
			got = zip_fread(entry, &databuffer[0], pixels*60 + 100000);
			if(got > pixels*60 + 10000)
			{
				printf("buffer still to small (read %ld/%ld). abort.\n", got, pixels*60 + 100000);
				exit(1);
			}
		}
		return got;
	}

	// times.txt: "id stamp [exposure_ms]" per line; on a count mismatch everything is zeroed
	void loadTimestamps(std::string timesFile)
	{
		std::ifstream in(timesFile.c_str());
		for(std::string line; std::getline(in, line);)
		{
			int id; double stamp; float exposure = 0;
			const int fields = sscanf(line.c_str(), "%d %lf %f", &id, &stamp, &exposure);
			if(fields < 2) continue;
			timestamps.push_back(stamp);
			exposures.push_back(fields == 3 ? exposure : 0);
		}
		if((int)exposures.size() != getNumImages())
		{
			printf("DatasetReader: Mismatch between number of images and number of timestamps / exposure times. Set all to zero.");
			timestamps.assign(getNumImages(), 0.0);
			exposures.assign(getNumImages(), 0.0f);
		}
	}

#endif

	// calibration models on the host + one device context holding all four tables for the fused kernel
	void loadCalibration()
	{
		undistorter = new UndistorterFOV((path + "camera.txt").c_str());
		widthOrg = undistorter->getInputDims()[0];
		heightOrg = undistorter->getInputDims()[1];
		width = undistorter->getOutputDims()[0];
		height = undistorter->getOutputDims()[1];
		photoUndistorter = new PhotometricUndistorter(path + "pcalib.txt", path + "vignette.png", widthOrg, heightOrg);
		if(mdc_ctx_create(UndistorterFOV::b200Device(), undistorter->b200Model(), photoUndistorter->b200Model(),
				&deviceContext) != MDC_OK)
		{
			printf("DatasetReader: cannot create the B200 device context: %s\n", mdc_last_error());
			deviceContext = 0;
		}
	}

	std::vector<std::string> files;
	std::vector<double> timestamps;
	std::vector<float> exposures;
	int width, height;          // rectified size
	int widthOrg, heightOrg;    // raw size
	std::string path;
	bool isZipped;

	UndistorterFOV* undistorter;
	PhotometricUndistorter* photoUndistorter;
#ifdef MDC_NATIVE_SEQUENCE_READER
	mdc_seq* ziparchive;
#else
	zip_t* ziparchive;
#endif
	std::vector<char> databuffer;
	mdc_ctx* deviceContext;
};
