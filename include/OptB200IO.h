/* OptB200IO.h -- minimal readers/writers for the file formats the reference's example programs and fixtures use, so that the
 * parity/benchmark suite needs neither mLib, LodePNG nor OpenMesh (SURVEY.md section 8(f) rank 4 "harness/IO independence").
 * C ABI, host memory only, plain pointers and sizes; every buffer handed out is released with the matching OptIO_Free* call.
 * All functions return 0 on success and a negative code on failure (-1 cannot open, -2 malformed / unsupported, -3 bad argument);
 * OptIO_LastError() gives the message of the calling thread's last failure.
 *
 * What each one replaces in the reference tree:
 *   OptIO_ReadImageDump / WriteImageDump  examples/shape_from_shading/src/SimpleBuffer.cpp:12-49,:79-90 and API/src/im.t:28-53
 *                                         (4 x int32 header: width, height, channels, datatype {0 float32, 1 uint8}, then raw data)
 *   OptIO_ReadSFSParameters               examples/shape_from_shading/src/TerraSolverParameters.h:7-45 (160-byte struct dump)
 *   OptIO_ReadPNG / WritePNG              LodePNG::load / save as used by examples/image_warping/src/main.cpp:60-61 (RGBA8 out)
 *   OptIO_ReadMesh / WritePLY             OpenMesh::IO::read_mesh (examples/arap_mesh_deformation/src/main.cpp:51,
 *                                         examples/robust_nonrigid_alignment/src/main.cpp) for .ply (ascii, binary_little_endian),
 *                                         .obj and .off; polygons are kept as index lists
 *   OptIO_MeshNeighbourEdges              the vertex-vertex walk + createGraphFromNeighborLists, examples/arap_mesh_deformation/
 *                                         src/CombinedSolver.h:111-148 and examples/shared/OptGraph.h:64-75 (directed edge per
 *                                         (vertex, neighbour), grouped by head vertex, neighbours in ascending index order)
 *   OptIO_ReadMarkers                     examples/arap_mesh_deformation/src/LandMarkSet.h:16-49 (.mrk: n, then x y z radius vertex)
 *   OptIO_ReadPairConstraints             loadConstraints, examples/image_warping/src/main.cpp:4-27 (n, then 4 ints per line)
 */
#ifndef OPTB200_IO_H
#define OPTB200_IO_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct OptIO_Image {
    int width, height, channels;
    int type;            /* 0 = float32, 1 = uint8 (the .imagedump datatype codes) */
    void* data;          /* width * height * channels elements, x fastest, channel-interleaved */
} OptIO_Image;

typedef struct OptIO_Mesh {
    int nVertices, nFaces;
    float* vertices;     /* 3 per vertex */
    int* faceStart;      /* nFaces + 1 offsets into faceIndices */
    int* faceIndices;
} OptIO_Mesh;

const char* OptIO_LastError(void);

int OptIO_ReadImageDump(const char* path, int clampInfinity, OptIO_Image* out);   /* clampInfinity: +inf -> FLT_MAX, -inf -> -10000 (float images) */
int OptIO_WriteImageDump(const char* path, const OptIO_Image* img);
int OptIO_ReadPNG(const char* path, OptIO_Image* out);                             /* always 4 channels uint8 (RGBA), any non-interlaced 8/16-bit PNG */
int OptIO_WritePNG(const char* path, const OptIO_Image* img);                      /* uint8, 1/3/4 channels */
void OptIO_FreeImage(OptIO_Image* img);

int OptIO_ReadSFSParameters(const char* path, float out40[40]);                    /* the 36 floats of the struct, then its 3 uints as floats; out40[39] = 0 */

int OptIO_ReadMesh(const char* path, OptIO_Mesh* out);                             /* by extension: .ply .obj .off */
int OptIO_WritePLY(const char* path, const OptIO_Mesh* mesh);                      /* ascii */
void OptIO_FreeMesh(OptIO_Mesh* mesh);
/* directed edges (head[i] -> tail[i]); *nEdges = 2 x undirected mesh edges.  Buffers are malloc'ed: release with OptIO_FreeInts. */
int OptIO_MeshNeighbourEdges(const OptIO_Mesh* mesh, int** head, int** tail, int* nEdges);
void OptIO_FreeInts(int* p);

/* .mrk: returns the count; fills at most maxMarkers entries of position (3 floats each), radius, vertex index */
int OptIO_ReadMarkers(const char* path, int maxMarkers, float* positions, float* radii, int* vertexIndices);
/* .constraints: returns the count; fills at most maxConstraints rows of 4 ints (x y targetX targetY) */
int OptIO_ReadPairConstraints(const char* path, int maxConstraints, int* rows);

#ifdef __cplusplus
}
#endif
#endif
