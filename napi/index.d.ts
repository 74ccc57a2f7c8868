// Type surface phaneron's sources use from 'nodencl' (SURVEY.md section 8b), as implemented by napi/phaneron_napi.cc.
/// <reference types="node" />
export type KernelParams = { [key: string]: unknown }
export type ImageDims = { width: number; height: number }
export type RunTimings = { dataToKernel: number; kernelExec: number; totalTime: number }
export interface OpenCLProgram {
	readonly name: string
	readonly op: string
	readonly width: number
	readonly height: number
}
export interface OpenCLBuffer extends Buffer {
	readonly numBytes: number
	readonly owner: string
	timestamp: number
	loadstamp: number
	creationTime: number
	hostAccess(mode?: 'none' | 'readonly' | 'writeonly', queue?: number, src?: Buffer): Promise<void>
	addRef(): void
	release(): void
	refs(): number
}
export interface RouteComm {
	begin(): void
	send(buf: OpenCLBuffer, peer: number): void
	recv(buf: OpenCLBuffer, peer: number): void
	end(): void
	wait(queue?: number, age?: number): void
	sync(): Promise<void>
	info(): { rank: number; world: number; bytesSent: number; bytesReceived: number }
	close(): void
}
export class clContext {
	constructor(options: { platformIndex: number; deviceIndex: number; overlapping?: boolean; deferred?: boolean })
	readonly queue: { load: number; process: number; unload: number }
	initialise(): Promise<void>
	getPlatformInfo(): { vendor: string; devices: { type: string; name: string }[] }
	createBuffer(numBytes: number, bufDir: 'readonly' | 'writeonly' | 'readwrite', bufType: 'none' | 'coarse' | 'fine', imageDims?: ImageDims, owner?: string): Promise<OpenCLBuffer>
	createProgram(
		source: string,
		options: { name: string; globalWorkItems: number | Uint32Array; workItemsPerGroup?: number; op?: string; width?: number; height?: number }
	): Promise<OpenCLProgram>
	runProgram(program: OpenCLProgram, params: KernelParams, queue: number): Promise<RunTimings>
	waitFinish(queue?: number): Promise<void>
	logBuffers(): unknown
	stats(): { [key: string]: number }
	setFlags(flags: number): void
	createComm(rank: number, world: number, uniqueId: Buffer): Promise<RouteComm>
	close(): void
	static uniqueId(): Buffer
	static routeCopyPeer(src: OpenCLBuffer, dst: OpenCLBuffer): Promise<void>
}
export function gamma2linearLUT(colSpec: string): Float32Array
export function linear2gammaLUT(colSpec: string): Float32Array
export const version: string
