// arks_nccl.h -- NCCL bound at run time (dlopen), so that libarks_b200.so has no link-time dependency on it:
// a single-GPU run never touches NCCL, and a process that already carries an NCCL (torch's, same SONAME) shares it.
// Only the types come from <nccl.h>.
#pragma once
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <string>

namespace arks {

struct NcclApi
{
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	std::string error; // why loading failed
	bool ok = false;
};

inline const NcclApi& nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		void* lib = nullptr;
		for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
			lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
			if (lib)
				break;
		}
		if (!lib) {
			api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
			return;
		}
		bool all = true;
		auto sym = [&](const char* n) {
			void* p = dlsym(lib, n);
			if (!p) {
				all = false;
				api.error = std::string("libnccl lacks ") + n;
			}
			return p;
		};
		api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
		api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
		api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
		api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
		api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
		api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
		api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
		api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
		api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
		api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
		api.ok = all;
	});
	return api;
}

} // namespace arks
