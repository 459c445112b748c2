/*
 * jpegb200_host.h -- C view of the host-side marker/header walk (libjpegb200_host.so).
 *
 * In the C# integration this work is done by the reference's own JpegDecoder.Identify() /
 * marker loop (JpegDecoder.cs:75-162, :509-617) and the result is marshalled into
 * jb_image_desc.  No .NET toolchain exists in the build image, so the same walk is provided
 * here in C++ (jpeglibrary_b200/host/jpeg_host.cpp) for the Python mirror of the reference API
 * (jpeglibrary_b200/api.py) that the tests and bench.py use.  It stands for Identify() followed by
 * Decode(): sequential and lossless frames get the restart interval JpegDecoder holds at the SOF.
 * It is host code only: no decode arithmetic lives here.
 */
#ifndef JPEGB200_HOST_H
#define JPEGB200_HOST_H

#include "jpegb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jbh_parsed jbh_parsed;

/* Walks the markers of one JPEG stream (SOI .. EOI) like JpegDecoder.Decode's loop and builds
   the descriptor the GPU path consumes.  `data` must stay alive while the descriptor is used.
   Returns JB_OK or JB_ERR_*; on failure *out is NULL and jbh_last_parse_error() has the text. */
JB_API int jbh_parse(const uint8_t *data, uint64_t length, jbh_parsed **out);
/* The same for an ABBREVIATED stream: `tables` is what the caller would hand to JpegDecoder.LoadTables
   (JpegDecoder.cs:313-360) before SetInput -- e.g. the JPEGTables field of a TIFF file.  Its DHT, DQT and DRI segments
   are the state the marker loop of `data` starts from (SOI / RSTn passed over, other segments skipped, EOI or the end
   of the data ends the walk); the stream's own segments replace them as they are met.  tables == NULL: jbh_parse. */
JB_API int jbh_parse_with_tables(const uint8_t *tables, uint64_t tables_length, const uint8_t *data, uint64_t length,
                                 jbh_parsed **out);
/* What JpegDecoder.LoadTables(tables) itself raises (a damaged DHT / DQT / DRI segment): JB_OK or JB_ERR_INVALID_DATA
   with the text in jbh_last_parse_error().  *used = the bytes the walk took: everything before an EOI or before bytes
   without a further marker, so that the streams of successive LoadTables calls can be kept back to back. */
JB_API int jbh_check_tables(const uint8_t *tables, uint64_t tables_length, uint64_t *used);
JB_API const jb_image_desc *jbh_desc(const jbh_parsed *p);
/* JpegDecoder.Identify(): bytes consumed up to and including EOI (MetadataIdentifyTests). */
JB_API uint64_t jbh_consumed(const jbh_parsed *p);
JB_API int jbh_sof_marker(const jbh_parsed *p); /* 0xC0.. */
JB_API void jbh_free(jbh_parsed *p);
JB_API const char *jbh_last_parse_error(void);

/* Parse `count` streams on `threads` host threads (bench / batch facade). out[i] receives the
   parsed object or NULL; returns the number of failures. */
JB_API int jbh_parse_batch(const uint8_t *const *data, const uint64_t *length, int count, int threads,
                           jbh_parsed **out);
/* `count` abbreviated streams that share one tables stream (the strips / tiles of one TIFF file). */
JB_API int jbh_parse_batch_with_tables(const uint8_t *tables, uint64_t tables_length, const uint8_t *const *data,
                                       const uint64_t *length, int count, int threads, jbh_parsed **out);
/* Copy the descriptors of parsed objects into a contiguous array for jb_decode_batch_create. */
JB_API int jbh_collect_descs(jbh_parsed *const *parsed, int count, jb_image_desc *descs);

#ifdef __cplusplus
}
#endif
#endif
