// Test-infrastructure stand-in for the libzip declarations BenchmarkDatasetReader.h names
// (SURVEY.md §8c).  Every call reports failure: the tests use an images/ folder, never a zip.
#pragma once
struct zip_t; struct zip_file_t;
#define ZIP_RDONLY 16
#define ZIP_FL_ENC_STRICT 128
inline zip_t* zip_open(const char*, int, int* err) { if (err) *err = 9; return 0; }
inline long zip_get_num_entries(zip_t*, int) { return 0; }
inline const char* zip_get_name(zip_t*, long, int) { return ""; }
inline zip_file_t* zip_fopen(zip_t*, const char*, int) { return 0; }
inline long zip_fread(zip_file_t*, void*, long) { return -1; }
inline int zip_close(zip_t*) { return 0; }
