// DatasetReader with MDC_NATIVE_SEQUENCE_READER: folder / zip / times.txt / frame decode through libmdc_b200 (no libzip, no
// cv::imread).  Dumps what the reader sees so that the Python test can compare: argv[1] = sequence folder, argv[2] = output file.
// Runs without a GPU (the device context fails to come up, which only getImage needs).
#define MDC_NATIVE_SEQUENCE_READER
#include "BenchmarkDatasetReader.h"

int main(int argc, char** argv)
{
	if(argc < 3) return 2;
	DatasetReader reader(argv[1]);
	FILE* out = fopen(argv[2], "wb");
	if(!out) return 3;
	const int n = reader.getNumImages();
	fwrite(&n, sizeof n, 1, out);
	for(int i = 0; i < n; i++)
	{
		const double stamp = reader.getTimestamp(i);
		const float exposure = reader.getExposure(i);
		cv::Mat raw = reader.getImageRaw_internal(i);
		const int dims[2] = { raw.rows, raw.cols };
		fwrite(&stamp, sizeof stamp, 1, out);
		fwrite(&exposure, sizeof exposure, 1, out);
		fwrite(dims, sizeof dims, 1, out);
		if(raw.rows * raw.cols > 0) fwrite(raw.data, 1, (size_t)raw.rows * raw.cols, out);
	}
	fclose(out);
	return 0;
}
